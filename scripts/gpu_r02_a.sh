#!/bin/bash
# round 2, first GPU call: the whole GPU test suite on the new ABI, the default bench lines, cooperative-launch A/B of
# the quantiser, comparators.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -15 gpurun_out/r02a_pytest.log
for coop in 0 1 2; do
  QA_RING_COOP=$coop timeout 120 python scripts/quant_time.py > gpurun_out/r02a_quant_coop$coop.txt 2>&1; tail -4 gpurun_out/r02a_quant_coop$coop.txt
done
timeout 600 python bench.py --steps 300 > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err; tail -3 gpurun_out/r02a_bench_c2.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_c2_driver.json 2> gpurun_out/r02a_bench_c2_driver.err
timeout 600 python bench.py --workload C3_llama --steps 200 --no-cpu-baseline > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err
python - <<'PY'
import json
for n in ("c2", "c2_driver", "c3"):
    try:
        d = json.load(open(f"gpurun_out/r02a_bench_{n}.json"))
        r = d["roofline"]
        print(n, "step", round(d["ms_per_step"] * 1e3, 1), "us", round(d["value"]), "TF/s | kernel", round(r["attn_kernel_ms"] * 1e3, 1), "us",
              round(r["achieved"]), "frac", round(r["frac"], 3), "| quant", round(d["quantiser"]["ms"] * 1e3, 1), "us frac", round(d["quantiser"]["frac"], 3),
              "| host_us", d.get("host_us_per_step"), "| e2e", round(d["e2e"]["value"], 1), "| acc", d.get("accuracy"), "| cmp", d.get("comparators"))
    except Exception as e:
        print(n, "FAILED", e)
PY
bash scripts/cutedsl_fmha_bar.sh
