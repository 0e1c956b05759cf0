#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py tests/test_ring_gpu.py -m gpu -q -x > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -8 gpurun_out/r02f_pytest.log
for rep in 1 2; do
  QA_NATIVE_LIB=$L/libqattn_sm100_noqt.so timeout 300 python scripts/ab_kernels.py noqt 2>&1 | tail -1
  timeout 300 python scripts/ab_kernels.py qtmem 2>&1 | tail -1
done | tee gpurun_out/r02f_ab.txt
