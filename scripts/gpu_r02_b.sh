#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -25 gpurun_out/r02b_pytest.log
for pdl in 1 0; do QA_PDL=$pdl timeout 120 python scripts/quant_probe.py 2>&1 | tail -2; done | tee gpurun_out/r02b_quant_probe.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:quant --csv --log-file gpurun_out/r02b_quant_launches.csv python scripts/quant_time.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r02b_quant_launches.csv")) if len(r) > 10 and r[0].isdigit()]
import collections
d = collections.defaultdict(list)
for r in rows:
    d[(r[4][:60], r[7] if len(r) > 7 else "")].append(float(r[-1].replace(",", "")))
for k, v in d.items():
    print(k, len(v), "launches, median", sorted(v)[len(v) // 2], "ns")
PY
bash scripts/cutedsl_fmha_bar.sh
