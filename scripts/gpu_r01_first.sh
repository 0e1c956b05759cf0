set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --workload C3_llama --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -o gpurun_out/attn_r01 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
