#!/bin/bash
# ablation A/B (timing only, the ablated kernels compute wrong results): how the step time responds to fewer MMAs / fewer exponentials
# variants first (here, no GPU): for v in "qk2 -DQA_ABL_QK=2" "pv2 -DQA_ABL_PV=2" "qk2pv2 -DQA_ABL_QK=2 -DQA_ABL_PV=2" "exp16 -DQA_ABL_EXP=16" "exp1 -DQA_ABL_EXP=1"; do set -- $v; n=$1; shift; bash scripts/build_variant.sh abl_$n "$@"; done
mkdir -p gpurun_out
L=quantumattention_b200
out=gpurun_out/r02_ablation.txt
: > $out
AB_SHAPES=C2,C3 AB_MODES=16bit,token python scripts/ab_kernels.py base_token >> $out 2>&1
for v in "" _abl_qk2 _abl_pv2 _abl_qk2pv2 _abl_exp16 _abl_exp1 ""; do
  QA_NATIVE_LIB=$PWD/$L/libqattn_sm100$v.so AB_SHAPES=C2,d64 AB_MODES=16bit,fp8 python scripts/ab_kernels.py "base$v" >> $out 2>&1
done
cat $out
