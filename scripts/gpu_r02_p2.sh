#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
export AB_SHAPES=C2,C3,d256,d64c
for rep in 1 2; do
  QA_NATIVE_LIB=$L/libqattn_sm100_old.so timeout 300 python scripts/ab_kernels.py old 2>&1 | tail -1
  QA_PERSIST=0 timeout 300 python scripts/ab_kernels.py plain 2>&1 | tail -1
  timeout 300 python scripts/ab_kernels.py persist 2>&1 | tail -1
done | tee gpurun_out/r02p2_ab.txt
