#!/bin/bash
# scripts/gpu_prof.sh VARIANT [workload]: ncu --set full capture of the attention kernel for one variant
v=$1; wl=${2:-C2_flux}
mkdir -p gpurun_out
QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_$v.so ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/prof_$v python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/prof_$v.log 2>&1
tail -2 gpurun_out/prof_$v.log | cut -c1-300
