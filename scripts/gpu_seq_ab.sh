#!/bin/bash
# C4 (one 75 600-token sequence over N GPUs): neighbour ring vs one all-gather (QA_SEQ_STRATEGY); NCCL parity test first
N=${1:-2}; T=${2:-s}
mkdir -p gpurun_out
if [ "$3" != "--no-tests" ]; then
  timeout 600 python -m pytest tests/test_ring_gpu.py -m gpu -x -q > gpurun_out/pytest_ring_$T.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ring_$T.log
  tail -4 gpurun_out/pytest_ring_$T.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for st in ring gather; do
  QA_SEQ_STRATEGY=$st timeout 600 $TR bench.py --gpus $N --workload C4_video --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/bench_c4_n${N}_${st}_$T.json 2> gpurun_out/bench_c4_n${N}_${st}_$T.err
  tail -3 gpurun_out/bench_c4_n${N}_${st}_$T.err
  python - gpurun_out/bench_c4_n${N}_${st}_$T.json $st <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "ms/step", round(d["ms_per_step"],3), "TFLOP/s", round(d["value"]), "attn launch ms", d["roofline"]["attn_kernel_ms"], d["clocks"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done
