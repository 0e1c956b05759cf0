#!/bin/bash
# External bar: NVIDIA's CuTe-DSL Blackwell FMHA example shipped inside the flashinfer wheel,
# run with Float8E4M3FN / Float16 / BFloat16 inputs on the C2 and C3 shapes - if its JIT works offline on the box.
# Not the reference and not product code: a comparator.  Output: gpurun_out/cutedsl_fmha_*.log
CUT=$(python - <<'PY'
import flashinfer, os
print(os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass", "examples", "python", "CuTeDSL", "blackwell"))
PY
)
mkdir -p gpurun_out
for cfg in "c2 1,4608,24,128 --" "c3 1,8192,32,128 --is_causal"; do
  set -- $cfg
  for dt in Float8E4M3FN BFloat16; do
    out=BFloat16
    timeout 240 python $CUT/fmha.py --in_dtype $dt --out_dtype $out --q_shape $2 --k_shape $2 $3 --is_persistent \
        --skip_ref_check --warmup_iterations 3 --iterations 10 > gpurun_out/cutedsl_fmha_$1_$dt.log 2>&1
    echo "cutedsl $1 $dt rc=$? : $(tail -2 gpurun_out/cutedsl_fmha_$1_$dt.log | tr '\n' ' ')"
  done
done
