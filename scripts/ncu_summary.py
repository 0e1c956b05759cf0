#!/usr/bin/env python
"""Print the metrics of interest from `ncu -i X.ncu-rep --page raw --csv` (run here, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
keys = sys.argv[2:] or ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor',
        'sm__inst_executed_pipe', 'smsp__inst_executed.sum', 'smsp__issue_active', 'sm__warps_active',
        'launch__registers', 'sm__cycles_elapsed.avg', 'sm__throughput', 'smsp__average_warp',
        'smsp__warps_issue_stalled', 'lts__t_bytes.sum', 'sm__cycles_active.avg', 'launch__grid', 'launch__block',
        'smsp__inst_executed_pipe', 'sm__inst_executed.avg.per_cycle', 'l1tex__data_bank', 'smsp__pcsamp_warps_issue']
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in keys):
            print(f"{h} [{u}] = {v}")
