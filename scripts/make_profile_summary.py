#!/usr/bin/env python
"""Turn gpurun_out/{launches,attn,quant}_<round>.* into the tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, json, os, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
lines = [f"# ncu summaries, round {R} (B200, `--set full --clock-control none`; per-launch, cold-cache, serialised)", ""]
traffic = {}
for name in ("attn", "quant"):
    rep = os.path.join(G, f"{name}_{R}.ncu-rep")
    if not os.path.exists(rep):
        continue
    hdr, units, rows = raw(rep)
    for r in rows:
        kn = r[hdr.index("Kernel Name")]
        lines += [f"## {kn[:110]}", "", "| metric | value | unit |", "|---|---|---|"]
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        for k in KEYS:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u[k]} |")
        lines.append("")
        if "dram__bytes_read.sum" in d:
            def tobytes(v, unit):
                v = float(v.replace(",", "")); return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]
            tb = tobytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) + tobytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
            traffic[kn.split("<")[0].split("::")[-1].replace("void ", "").strip()] = tb
open(os.path.join(P, f"{R}_ncu_summary.md"), "w").write("\n".join(lines))

# launch list -> per-kernel table + the raw csv
src = os.path.join(G, f"launches_{R}.csv")
if os.path.exists(src):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]; ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault((r[ki][:90], r[gi], r[bi]), []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu launch list of `python bench.py --steps 8 --warmup 3` (round {R}); durations are cold-cache and serialised: compare SHARES", "",
           "| kernel | grid | block | launches | mean us | share of listed time |", "|---|---|---|---|---|---|"]
    for (k, g, b), v in agg.items():
        out.append(f"| `{k}` | {g} | {b} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    open(os.path.join(P, f"{R}_launches.md"), "w").write("\n".join(out) + "\n")
    open(os.path.join(P, f"{R}_launches.csv"), "w").write(open(src).read())

tj = os.path.join(P, "roofline_traffic.json")
cur = json.load(open(tj)) if os.path.exists(tj) else {}
for k, v in traffic.items():
    if "attn_fwd" in k:
        # AttnCfg<D, PMODE, ...>: the P mode is the kernel's second template argument (0 fp8, 1 fp8_hilo, 2 16bit)
        pm = {"0": "fp8", "1": "fp8_hilo", "2": "16bit"}.get(os.environ.get("QA_SUMMARY_PMODE", "0"), "fp8")
        cur.setdefault("C2_flux", {})[pm] = v
    else:
        cur.setdefault("quantiser", {})[k] = v
cur["_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, from the --set full captures summarised beside this file"
json.dump(cur, open(tj, "w"), indent=1)
print("\n".join(lines[:40]))
