#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
QA_NATIVE_LIB=$L/libqattn_sm100_split.so timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py -m gpu -q -x > gpurun_out/r02d_pytest_split.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest_split.log
tail -8 gpurun_out/r02d_pytest_split.log
for rep in 1 2; do
  timeout 300 python scripts/ab_kernels.py base 2>&1 | tail -1
  QA_NATIVE_LIB=$L/libqattn_sm100_split.so timeout 300 python scripts/ab_kernels.py split 2>&1 | tail -1
done | tee gpurun_out/r02d_ab.txt
timeout 300 python scripts/cutedsl_fmha_bar.py > gpurun_out/r02d_cutedsl.log 2>&1; grep -E "^C[23]" gpurun_out/r02d_cutedsl.log
