#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_ring_gpu.py -m gpu -q -k "nccl" -s > gpurun_out/r02m2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m2d_pytest.log
grep -E "passed|failed|peer transport|Error" gpurun_out/r02m2d_pytest.log | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 4 > gpurun_out/r02m2d_bench_n$N.json 2> gpurun_out/r02m2d_bench_n$N.err; echo "bench rc=$?"; tail -2 gpurun_out/r02m2d_bench_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02m2d_bench_n$N.json"))["seq_sharded"]
print("N=$N:", d["ms_per_step"], "eff", d["strong_scaling_efficiency"], d["transport"], d["wire"]["gather_alone_ms"], d["wire"]["gather_alone_gbs"])
print(json.dumps(d["time_split"], indent=0)[:1200])
PY
