#!/usr/bin/env python
"""SASS opcode census of libqattn_sm100.so (runs here: cuobjdump needs no GPU) -> profiles/<round>_sass_census.md.
What proves a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UBLKCP = TMA, no HMMA / HGMMA (legacy mma.sync / Hopper wgmma)."""
import collections, os, re, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "quantumattention_b200", "libqattn_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
OPS = ["UTCQMMA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "MUFU.EX2", "F2FP", "FFMA2",
       "HMMA", "HGMMA", "QGMMA", "ACQBULK", "USETMAXREG"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("qa::", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                per[cur][o] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
cols = [o for o in OPS if tot[o]] + [o for o in ("HMMA", "HGMMA", "QGMMA") if not tot[o]]
out = [f"# SASS opcode census of `libqattn_sm100.so` (round {R}; `cuobjdump -sass`, architectures in the file: {', '.join(arch)})", "",
       "`UTCQMMA` / `UTCHMMA` = `tcgen05.mma` kind::f8f6f4 / kind::f16, `LDTM` / `STTM` = `tcgen05.ld` / `st`, `UTMALDG` / `UBLKCP` = TMA tensor / bulk "
       "copies, `SYNCS` = mbarrier ops, `UTCBAR` = `tcgen05.commit`; no legacy `HMMA` and no Hopper `*GMMA` anywhere.", "",
       "| kernel | instructions | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
for k, c in per.items():
    out.append(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[o]) for o in cols) + " |")
out.append(f"| **all {len(per)} kernels** | {tot['_total']} | " + " | ".join(f"**{tot[o]}**" for o in cols) + " |")
path = os.path.join(ROOT, "profiles", f"{R}_sass_census.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
print("wrote", path)
