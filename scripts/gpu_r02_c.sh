#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -6 gpurun_out/r02c_pytest.log
QA_NATIVE_LIB=$L/libqattn_sm100_split.so timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py tests/test_sweep_gpu.py tests/test_ring_gpu.py -m gpu -q --maxfail=30 > gpurun_out/r02c_pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02c_pytest_split.log
tail -12 gpurun_out/r02c_pytest_split.log
for rep in 1 2; do
  timeout 300 python scripts/ab_kernels.py base 2>&1 | tail -1
  QA_NATIVE_LIB=$L/libqattn_sm100_split.so timeout 300 python scripts/ab_kernels.py split 2>&1 | tail -1
done | tee gpurun_out/r02c_ab.txt
timeout 300 python scripts/cutedsl_fmha_bar.py 2>&1 | grep -E "^C[23]" | tee gpurun_out/r02c_cutedsl.txt
timeout 300 python bench.py --steps 300 --no-cpu-baseline --no-comparators --no-other-modes > gpurun_out/r02c_bench_c2.json 2> gpurun_out/r02c_bench_c2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_bench_c2.json")); print("quantiser", d["quantiser"]["ms"], d["quantiser"]["frac"], "step", d["ms_per_step"], "host_us", d["host_us_per_step"])
PY
