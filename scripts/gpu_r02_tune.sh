#!/bin/bash
# A/B of the softmax schedule knobs and the per-warp arrival on the final kernel (kernel alone, one box)
# variants first (here, no GPU): bash scripts/build_variant.sh tune_warparrive -DQA_WARP_ARRIVE=1; tune_loadq6 -DQA_LOADQ=6; tune_loadq10 -DQA_LOADQ=10;
# tune_ldw2 -DQA_LDWAITQ=2; tune_ldw6 -DQA_LDWAITQ=6; tune_pubq4 -DQA_PUBQ=4; tune_poly3 -DQA_POLY_P16=3; tune_poly1 -DQA_POLY_P16=1; tune_dec2 -DQA_DECIDEQ=2
mkdir -p gpurun_out
L=$PWD/quantumattention_b200
out=gpurun_out/r02_tune_ab.txt
: > $out
for v in "" _tune_warparrive _tune_loadq6 _tune_loadq10 _tune_ldw2 _tune_ldw6 _tune_pubq4 _tune_poly3 _tune_poly1 _tune_dec2 ""; do
  QA_NATIVE_LIB=$L/libqattn_sm100$v.so AB_SHAPES=C2,C3,d64 AB_MODES=16bit,fp8 python scripts/ab_kernels.py "base$v" >> $out 2>&1
done
QA_NATIVE_LIB=$L/libqattn_sm100_tune_warparrive.so timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_attention16_gpu.py -m gpu -q -x 2>&1 | tail -2 >> $out
cat $out
