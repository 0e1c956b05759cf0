#!/bin/bash
# gpurun_out/*_ROUND.* (scripts/gpu_r02_final1.sh, scripts/gpu_r02_drv.sh) -> the tracked artefacts under profiles/.  Runs here.
R=${1:-r02}
cd "$(dirname "$0")/.."
python scripts/make_profile_summary.py $R > /dev/null
python scripts/stall_table.py $R > /dev/null
for n in c1 c2 c2_fp8 c3 c4 ref; do cp gpurun_out/bench_${n}_$R.json profiles/${R}_bench_$n.json; done
# the driver's command: first of the repeats of scripts/gpu_r02_drv.sh when they exist
if [ -f gpurun_out/bench_c2_driver_rep1.json ]; then
  cp gpurun_out/bench_c2_driver_rep1.json profiles/${R}_bench_c2_driver.json
  python - <<'PY' > profiles/r02_bench_c2_driver_repeats.txt
import json
print("python bench.py --steps 20 --warmup 5 (the driver's N = 1 command), four runs in one gpurun call")
for i in (1, 2, 3, 4):
    d = json.load(open(f"gpurun_out/bench_c2_driver_rep{i}.json")); r = d["roofline"]
    print(f"run {i}: step {d['ms_per_step'] * 1e3:.1f} us, {d['value']:.0f} TFLOP/s; attention kernel {r['attn_kernel_ms'] * 1e3:.1f} us, "
          f"{r['achieved']:.0f} TFLOP/s, frac {r['frac']:.3f}; e2e {d['e2e']['value']:.1f}; SM clock {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")
PY
else
  cp gpurun_out/bench_c2_driver_$R.json profiles/${R}_bench_c2_driver.json
fi
cp gpurun_out/sweep_$R.md profiles/${R}_sweep.md
cp gpurun_out/kernels_$R.txt profiles/${R}_kernels_by_shape.txt
cp gpurun_out/quant_time_$R.txt profiles/${R}_quantiser_graph_timed.txt
grep -h CUTEDSL_JSON gpurun_out/cutedsl_$R.log | sed 's/^CUTEDSL_JSON *//' > profiles/${R}_cutedsl_fmha_bar.json
for t in memcheck racecheck synccheck; do cp gpurun_out/sanitizer_${t}_$R.txt profiles/${R}_sanitizer_$t.txt; done
python scripts/sass_census.py > profiles/${R}_sass_census.md 2>/dev/null || true
ls -la profiles | grep $R | wc -l
