#!/bin/bash
# Quick validation on the B200 box: GPU tests + C2/C3 bench lines into gpurun_out/
set -x
T=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$T.log
tail -5 gpurun_out/pytest_gpu_$T.log
timeout 600 python bench.py > gpurun_out/bench_c2_$T.json 2> gpurun_out/bench_c2_$T.err; tail -3 gpurun_out/bench_c2_$T.err
timeout 600 python bench.py --workload C3_llama --no-cpu-baseline > gpurun_out/bench_c3_$T.json 2> gpurun_out/bench_c3_$T.err
cat gpurun_out/bench_c2_$T.json gpurun_out/bench_c3_$T.json
