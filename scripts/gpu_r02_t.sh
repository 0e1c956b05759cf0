#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
export AB_SHAPES=C3,d64c,d256c AB_MODES=16bit,fp8
for rep in 1 2; do
  timeout 300 python scripts/ab_kernels.py base 2>&1 | tail -1
  QA_NATIVE_LIB=$L/libqattn_sm100_tail2.so timeout 300 python scripts/ab_kernels.py tail2 2>&1 | tail -1
done | tee gpurun_out/r02t_ab.txt
QA_NATIVE_LIB=$L/libqattn_sm100_tail2.so timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 50 --no-cpu-baseline --no-comparators --no-other-modes 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('host_us', d['host_us_per_step'])"
