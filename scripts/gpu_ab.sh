#!/bin/bash
# A/B the kernel variants built by scripts/build_variant.sh on the C2 workload; prints kernel ms / TFLOP/s per variant
mkdir -p gpurun_out
if [ "$1" == "--tests" ]; then
  shift
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
for v in "$@"; do
  wl=C2_flux
  QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_$v.so timeout 300 python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline --no-other-modes --e2e-steps 2 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{v}.json"))
    r=d["roofline"]; print(f"{v}: kernel {r['attn_kernel_ms']*1e3:.1f} us  {r['achieved']:.0f} TF/s  step {d['ms_per_step']*1e3:.1f} us {d['value']:.0f} TF/s clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(v, "FAILED", e, open(f"gpurun_out/ab_{v}.err").read()[-500:])
PY
done
