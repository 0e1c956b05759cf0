#!/bin/bash
# 2-GPU call: NCCL / peer-memory sequence sharding parity, two devices from one process, bench lines at N = 2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02m2_topo.txt 2>&1
timeout 600 python -m pytest tests/test_ring_gpu.py tests/test_graph_compile_gpu.py -m gpu -q -k "nccl or two_devices or single_rank or head_group" -s > gpurun_out/r02m2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02m2_pytest.log
grep -E "passed|failed|peer transport|Error" gpurun_out/r02m2_pytest.log | tail -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m2_bench_n2.json 2> gpurun_out/r02m2_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02m2_bench_n2.err
for tr in nccl peer; do
  QA_SEQ_TRANSPORT=$tr timeout 600 $TR bench.py --gpus 2 --workload C4_video --steps 8 --warmup 3 --no-seq-sharded --e2e-steps 2 > gpurun_out/r02m2_bench_c4_$tr.json 2> gpurun_out/r02m2_bench_c4_$tr.err; echo "c4 $tr rc=$?"; tail -2 gpurun_out/r02m2_bench_c4_$tr.err
done
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02m2_bench_n2.json"))
    print("N2 C2: value", round(d["value"]), "ms", d["ms_per_step"], "e2e", round(d["e2e"]["value"], 1), d.get("host_link"), d.get("host_binding"))
    print("seq_sharded:", json.dumps(d.get("seq_sharded"), indent=1)[:3000])
except Exception as e:
    print("n2 failed", e)
for tr in ("nccl", "peer"):
    try:
        d = json.load(open(f"gpurun_out/r02m2_bench_c4_{tr}.json"))
        print("C4", tr, "ms", round(d["ms_per_step"], 3), "TF", round(d["value"]), d["config"].get("seq_transport"))
    except Exception as e:
        print("c4", tr, "failed", e)
PY
