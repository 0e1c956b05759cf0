#!/bin/bash
# the driver's N=1 command, repeated: run-to-run spread of the 20-step line
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_driver_rep$i.json 2>/dev/null
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_c2_driver_rep$i.json"))
r = d["roofline"]
print("rep$i step", round(d["ms_per_step"] * 1e3, 1), "us", round(d["value"]), "TF | kernel", round(r["attn_kernel_ms"] * 1e3, 1), "frac", round(r["frac"], 3), "| e2e", round(d["e2e"]["value"], 1), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["samples"], d["clocks"]["reasons"])
PY
done
