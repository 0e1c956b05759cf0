#!/bin/bash
# the driver's scaling command at N GPUs -> gpurun_out/bench_c2_n$N_r02.json (C2 weak scaling + the C4 seq_sharded block)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c2_n${N}_r02.json 2> gpurun_out/bench_c2_n${N}_r02.err; echo "bench rc=$?"
if [ "$N" == "2" ]; then
  timeout 600 python -m pytest tests/test_ring_gpu.py tests/test_graph_compile_gpu.py -m gpu -q -k "nccl or two_devices" > gpurun_out/pytest_gpu_n2_r02.log 2>&1; tail -2 gpurun_out/pytest_gpu_n2_r02.log
  QA_SEQ_STRATEGY=ring timeout 600 $TR bench.py --gpus $N --workload C4_video --steps 8 --warmup 3 --no-seq-sharded --e2e-steps 2 > gpurun_out/bench_c4_n${N}_ring_r02.json 2>/dev/null
fi
python - <<PY
import json
d = json.load(open("gpurun_out/bench_c2_n${N}_r02.json"))
print("N=$N C2: value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d.get("host_link"))
s = d["seq_sharded"]
print("seq:", round(s["ms_per_step"], 3), "ms", round(s["value"]), "TF eff", s["strong_scaling_efficiency"], s["transport"], s["transport_error"], "one gpu", s["one_gpu_ms_same_run"], s["wire"]["gather_alone_ms"], s["wire"]["gather_alone_gbs"])
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in s["time_split"].items() if k != "note"})
print(s["accuracy"])
PY
