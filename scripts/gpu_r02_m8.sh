#!/bin/bash
# 8-GPU call: the driver's scaling command at N = 8 (C2 weak scaling + the C4 strong-scaling block), and C4 as the main
# workload with both transports
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02m8_topo.txt 2>&1
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02m8_bench_n$N.json 2> gpurun_out/r02m8_bench_n$N.err; echo "bench n$N rc=$?"; tail -3 gpurun_out/r02m8_bench_n$N.err
for tr in nccl peer; do
  QA_SEQ_TRANSPORT=$tr timeout 600 $TR bench.py --gpus $N --workload C4_video --steps 10 --warmup 3 --no-seq-sharded --e2e-steps 2 > gpurun_out/r02m8_bench_c4_$tr.json 2> gpurun_out/r02m8_bench_c4_$tr.err; echo "c4 $tr rc=$?"; tail -2 gpurun_out/r02m8_bench_c4_$tr.err
done
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r02m8_bench_n$N.json"))
    print("N$N C2: value", round(d["value"]), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), d.get("host_link"), d.get("host_binding")[:2])
    print("seq_sharded:", json.dumps(d.get("seq_sharded"), indent=1)[:3500])
except Exception as e:
    print("failed", e)
for tr in ("nccl", "peer"):
    try:
        d = json.load(open(f"gpurun_out/r02m8_bench_c4_{tr}.json"))
        print("C4", tr, "ms", round(d["ms_per_step"], 3), "TF", round(d["value"]))
    except Exception as e:
        print("c4", tr, "failed", e)
PY
