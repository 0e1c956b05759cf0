#!/bin/bash
# round-2 artefacts on one B200: GPU tests, bench lines (C2 headline + driver-sized run, C3, C4, C1, fp8 mode, reference
# arm), kernel tables, sweep, comparators, ncu launch list + full captures.  Everything lands in gpurun_out/.
R=r02
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/smi_$R.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$R.log
tail -3 gpurun_out/pytest_gpu_$R.log
timeout 900 python bench.py > gpurun_out/bench_c2_$R.json 2> gpurun_out/bench_c2_$R.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2_driver_$R.json 2> gpurun_out/bench_c2_driver_$R.err
timeout 900 python bench.py --workload C3_llama > gpurun_out/bench_c3_$R.json 2> gpurun_out/bench_c3_$R.err
timeout 900 python bench.py --workload C4_video --no-cpu-baseline > gpurun_out/bench_c4_$R.json 2> gpurun_out/bench_c4_$R.err
timeout 600 python bench.py --workload C1 --no-cpu-baseline --no-comparators > gpurun_out/bench_c1_$R.json 2> gpurun_out/bench_c1_$R.err
timeout 600 python bench.py --pv-mode fp8 --no-cpu-baseline --no-comparators > gpurun_out/bench_c2_fp8_$R.json 2> gpurun_out/bench_c2_fp8_$R.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err
AB_MODES=16bit,fp8,hilo,token timeout 600 python scripts/ab_kernels.py $R > gpurun_out/kernels_$R.txt 2>&1
timeout 300 python scripts/quant_time.py > gpurun_out/quant_time_$R.txt 2>&1
timeout 400 python scripts/cutedsl_fmha_bar.py > gpurun_out/cutedsl_$R.log 2>&1
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitizer_${tool}_$R.txt 2>&1; done
QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_synccheck.so timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitizer_synccheck_$R.txt 2>&1
timeout 900 python scripts/sweep.py > gpurun_out/sweep_$R.md 2> gpurun_out/sweep_$R.err
BENCH="python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-comparators --no-other-modes"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$R.csv $BENCH > gpurun_out/ncu_list_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 3 -c 1 -f -o gpurun_out/attn_$R $BENCH > gpurun_out/ncu_attn_$R.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quant_head -s 6 -c 1 -f -o gpurun_out/quant_$R $BENCH > gpurun_out/ncu_quant_$R.log 2>&1
python - <<'PY'
import json
for n in ("c2", "c2_driver", "c3", "c4", "c1", "c2_fp8"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}_r02.json"))
        r = d["roofline"]
        print(n, "step", round(d["ms_per_step"] * 1e3, 1), "us", round(d["value"]), "TF/s | kernel", round(r["attn_kernel_ms"] * 1e3, 1), "us",
              round(r["achieved"]), "frac", round(r["frac"], 3), "| quant", round(d["quantiser"]["ms"] * 1e3, 1), "us frac", round(d["quantiser"]["frac"], 3),
              "| host_us", round(d.get("host_us_per_step") or 0, 1), "| e2e", round(d["e2e"]["value"], 1), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "FAILED", e)
PY
