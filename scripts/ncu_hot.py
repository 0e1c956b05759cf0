"""scripts/ncu_hot.py REPORT.csv [N]: hottest SASS instructions of an `ncu --page source --csv` export, with their stall mix."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("total samples", tot, " instructions", len(body))
print("stall mix:", {k[6:]: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.005})
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
top = sorted(body, key=lambda r: -int(r[col["# Samples"]] or 0))[:N]
for r in sorted(top, key=lambda r: r[col["Address"]]):
    s = int(r[col["# Samples"]] or 0)
    mix = sorted(((int(r[col[k]] or 0), k[6:]) for k in stalls), reverse=True)[:3]
    print(f"{r[col['Address']][-5:]} {100*s/tot:5.2f}% ex={r[col['Instructions Executed']]:>9} {r[col['Source']][:70]:70s} {[(n, c) for c, n in mix if c]}")
