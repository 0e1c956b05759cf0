#!/bin/bash
# one --set full capture (with source) of the attention kernel on the C2 workload -> gpurun_out/attn_$1.ncu-rep
T=${1:-x}; WL=${2:-C2_flux}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_$T python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-other-modes --e2e-steps 1 > gpurun_out/ncu_attn_$T.log 2>&1
tail -3 gpurun_out/ncu_attn_$T.log
