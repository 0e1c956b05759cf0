#!/bin/bash
# Final check after making "16bit" (the reference's P/V numerics) the default mode: GPU suite, smoke, bench lines, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f16.log
tail -4 gpurun_out/pytest_gpu_f16.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_c2_f16.json 2> gpurun_out/bench_c2_f16.err; tail -2 gpurun_out/bench_c2_f16.err
timeout 300 python bench.py --workload C3_llama --no-cpu-baseline > gpurun_out/bench_c3_f16.json 2> gpurun_out/bench_c3_f16.err
cut -c1-300 gpurun_out/bench_c2_f16.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_f16 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-modes --e2e-steps 1 > gpurun_out/ncu_attn_f16.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_f16.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-other-modes --e2e-steps 1 > gpurun_out/ncu_list_f16.log 2>&1
ls -la gpurun_out/*f16*
