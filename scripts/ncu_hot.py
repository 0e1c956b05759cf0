#!/usr/bin/env python
"""Top stall sites from `ncu -i X.ncu-rep --page source --csv` (SASS view)."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "total inst", sum(int(r[iex]) for r in data))
opc = collections.Counter()
for r in data:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    opc[op.split(".")[0]] += int(r[iex])
print("inst mix:", opc.most_common(25))
agg = collections.Counter()
for r in data:
    for i in stalls: agg[hdr[i]] += int(r[i])
print("stall totals:", agg.most_common(12))
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][isamp]))[:topn]:
    st = sorted(((int(r[i]), hdr[i][6:]) for i in stalls if int(r[i])), reverse=True)[:3]
    print(f"{idx:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>8} {r[isrc].strip()[:70]:70s} {st}")
