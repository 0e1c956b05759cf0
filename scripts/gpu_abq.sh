#!/bin/bash
# A/B quantiser variants built by scripts/build_variant.sh: quantiser leg (memset + kernel) of the C2 / C3 step
mkdir -p gpurun_out
for v in "$@"; do
  for wl in C2_flux C3_llama; do
    QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_$v.so timeout 300 python bench.py --workload $wl --steps 200 --warmup 20 --no-cpu-baseline --no-other-modes --e2e-steps 2 > gpurun_out/abq_$v.json 2> gpurun_out/abq_$v.err
    python - "$v" "$wl" <<'PY'
import json,sys
v,wl=sys.argv[1:3]
try:
    d=json.load(open(f"gpurun_out/abq_{v}.json")); q=d["quantiser"]
    print(f"{v} {wl}: quant {q['ms']*1e3:.1f} us {q['achieved']:.0f} GB/s  step {d['ms_per_step']*1e3:.1f} us")
except Exception as e:
    print(v, wl, "FAILED", e, open(f"gpurun_out/abq_{v}.err").read()[-300:])
PY
  done
done
