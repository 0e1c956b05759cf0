#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py tests/test_ring_gpu.py tests/test_sweep_gpu.py -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -5 gpurun_out/r02p_pytest.log
for rep in 1 2; do
  QA_PERSIST=0 timeout 300 python scripts/ab_kernels.py plain 2>&1 | tail -1
  QA_NATIVE_LIB=$L/libqattn_sm100_noqt.so timeout 300 python scripts/ab_kernels.py old 2>&1 | tail -1
  timeout 300 python scripts/ab_kernels.py persist 2>&1 | tail -1
done | tee gpurun_out/r02p_ab.txt
