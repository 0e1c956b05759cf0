#!/bin/bash
# Round artefacts on the B200 box: GPU tests, bench lines (C2 headline, C3), the ncu launch list of the bench command and
# one --set full capture of the attention kernel and of the quantiser.  Everything lands in gpurun_out/.
set -x
R=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_$R.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$R.log
python bench.py > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err
python bench.py --workload C3_llama --no-cpu-baseline > gpurun_out/bench_c3_$R.json 2> gpurun_out/bench_c3_$R.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_$R python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_attn_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quant_head\|amax_head -s 6 -c 2 -f -o gpurun_out/quant_$R python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_quant_$R.log 2>&1
ls -la gpurun_out | tail -20
