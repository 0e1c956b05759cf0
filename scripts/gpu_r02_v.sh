#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
export AB_MODES=16bit AB_SHAPES=C2,C3,d64,d256,c4s
for rep in 1 2; do
  QA_NATIVE_LIB=$L/libqattn_sm100_prev.so timeout 300 python scripts/ab_kernels.py prev 2>&1 | tail -1
  timeout 300 python scripts/ab_kernels.py pf16 2>&1 | tail -1
done | tee gpurun_out/r02v_ab.txt
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py tests/test_graph_compile_gpu.py -m gpu -q -x 2>&1 | tail -2
