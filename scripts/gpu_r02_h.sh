#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02h_pytest.log
tail -6 gpurun_out/r02h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02h_bench_c2_driver.json 2> gpurun_out/r02h_bench_c2_driver.err
timeout 600 python bench.py --workload C3_llama --steps 200 --no-cpu-baseline > gpurun_out/r02h_bench_c3.json 2> gpurun_out/r02h_bench_c3.err
timeout 600 python bench.py --workload C4_video --steps 10 --no-cpu-baseline --no-other-modes > gpurun_out/r02h_bench_c4.json 2> gpurun_out/r02h_bench_c4.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02h_bench_ref.json 2> gpurun_out/r02h_bench_ref.err
python - <<'PY'
import json
for n in ("c2_driver", "c3", "c4"):
    try:
        d = json.load(open(f"gpurun_out/r02h_bench_{n}.json"))
        r = d["roofline"]
        print(n, "step", round(d["ms_per_step"] * 1e3, 1), "us", round(d["value"]), "TF/s | kernel", round(r["attn_kernel_ms"] * 1e3, 1), "us",
              round(r["achieved"]), "frac", round(r["frac"], 3), "samples", r["kernel_samples"], "| quant", round(d["quantiser"]["ms"] * 1e3, 1), "us frac", round(d["quantiser"]["frac"], 3),
              "| host_us", d.get("host_us_per_step"), "| e2e", round(d["e2e"]["value"], 1), "| clocks", d["clocks"])
    except Exception as e:
        print(n, "FAILED", e)
a = json.load(open("gpurun_out/r02h_bench_c2_driver.json")); b = json.load(open("gpurun_out/r02h_bench_ref.json"))
print("same config:", a["config"] == b["config"], [k for k in a["config"] if a["config"].get(k) != b["config"].get(k)])
PY
