#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
for ns in 4; do
QA_PEER_STREAMS=$ns timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --e2e-steps 4 > gpurun_out/r02m2b_bench_n${N}_s$ns.json 2> gpurun_out/r02m2b_bench_n${N}_s$ns.err; echo "bench rc=$?"; tail -2 gpurun_out/r02m2b_bench_n${N}_s$ns.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02m2b_bench_n${N}_s$ns.json"))["seq_sharded"]
print("streams $ns:", d["ms_per_step"], d["strong_scaling_efficiency"], d["transport"], d["wire"]["gather_alone_ms"], d["wire"]["gather_alone_gbs"])
print(json.dumps(d["time_split"], indent=0)[:1500])
PY
done
