#!/bin/bash
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
timeout 60 scripts/ubench/ldtm_rate | tee gpurun_out/r02g_ldtm.txt
export AB_SHAPES=C2,C3,d64
QA_NATIVE_LIB=$L/libqattn_sm100_noqt.so timeout 300 python scripts/ab_kernels.py noqt 2>&1 | tail -1
QA_NATIVE_LIB=$L/libqattn_sm100_pb1.so timeout 300 python scripts/ab_kernels.py pb1 2>&1 | tail -1
timeout 300 python scripts/ab_kernels.py qtmem 2>&1 | tail -1
timeout 400 python scripts/cutedsl_fmha_bar.py > gpurun_out/r02g_cutedsl.log 2>&1; grep -E "^C[23]|Error" gpurun_out/r02g_cutedsl.log | tail
