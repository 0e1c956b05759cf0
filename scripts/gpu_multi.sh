#!/bin/bash
# Multi-GPU checks on an N-GPU box: NCCL ring test, weak-scaling bench (C2), strong-scaling ring bench (C4_video)
N=${1:-2}; T=${2:-m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/smi_multi_$T.txt
timeout 600 python -m pytest tests/test_ring_gpu.py -m gpu -x -q > gpurun_out/pytest_ring_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ring_$T.log
tail -4 gpurun_out/pytest_ring_$T.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 300 --warmup 10 --no-other-modes > gpurun_out/bench_c2_n${N}_$T.json 2> gpurun_out/bench_c2_n${N}_$T.err
tail -2 gpurun_out/bench_c2_n${N}_$T.err; cat gpurun_out/bench_c2_n${N}_$T.json
timeout 900 $TR bench.py --gpus $N --workload C4_video --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/bench_c4_n${N}_$T.json 2> gpurun_out/bench_c4_n${N}_$T.err
tail -5 gpurun_out/bench_c4_n${N}_$T.err; cat gpurun_out/bench_c4_n${N}_$T.json
