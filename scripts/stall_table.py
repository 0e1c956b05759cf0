#!/usr/bin/env python
"""scripts/stall_table.py ROUND: source-level stall table of the attention kernel's softmax main loop from
gpurun_out/attn_ROUND.ncu-rep (`ncu --set full --import-source on`) -> profiles/ROUND_attn_stalls.md.  Runs here."""
import collections, csv, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"attn_{R}.ncu-rep")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
S = "# Samples"
tot = sum(int(r[col[S]] or 0) for r in body)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ex = collections.Counter(int(r[col["Instructions Executed"]] or 0) for r in body if "MUFU.EX2" in r[col["Source"]])
loop_ex = ex.most_common(1)[0][0]
loop = [r for r in body if int(r[col["Instructions Executed"]] or 0) == loop_ex]
lt = sum(int(r[col[S]] or 0) for r in loop)
agg = {s: sum(int(r[col[s]] or 0) for r in loop) for s in stalls}
kname = rows[hi - 1][0] if hi > 0 and rows[hi - 1] else "attn_fwd_kernel"
out = [f"# Source-level stall sampling of the attention kernel on C2, default mode (round {R})", "",
       f"From `ncu --set full --import-source on` (`gpurun_out/attn_{R}.ncu-rep`, `--page source --csv`), restricted to the softmax main "
       f"loop (the {len(loop)} SASS instructions executed {loop_ex} times each = 8 softmax warps x 432 CTAs x 35 trips of two 64-key "
       f"steps): {lt} of {tot} warp samples.", "", "| warp state in the main loop | share |", "|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v > lt * 0.01:
        out.append(f"| {k[6:]} | {100 * v / lt:.1f}% |")
ops = collections.defaultdict(lambda: [0, 0])
for r in loop:
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    op = m.group(1) if m else "?"
    op = re.sub(r"\.(F32|PACK_AB.*|SATFINITE.*|FTZ|NAN|RN|U32|X4|WIDE.*)$", "", op)
    ops[op][0] += 1
    ops[op][1] += int(r[col[S]] or 0)
out += ["", "| opcode | instructions in the loop body (2 steps) | share of samples | samples per instruction |", "|---|---|---|---|"]
for op, (n, s_) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:14]:
    out.append(f"| `{op}` | {n} | {100 * s_ / lt:.1f}% | {s_ / n:.1f} |")
note = os.path.join(ROOT, "profiles", f"{R}_attn_stalls_note.md")
if os.path.exists(note):
    out += ["", open(note).read().strip()]
open(os.path.join(ROOT, "profiles", f"{R}_attn_stalls.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
