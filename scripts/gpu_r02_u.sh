#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_quantize_gpu.py -m gpu -q -k two_streams 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02u_pytest.log
tail -3 gpurun_out/r02u_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; tail -2 gpurun_out/r02u_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02u_bench.json"))
print("step", d["ms_per_step"], "value", d["value"], "clocks", d["clocks"], "host_us", d["host_us_per_step"], "quant", d["quantiser"]["ms"], d["quantiser"]["traffic"], "launches", d["gpu_launches"])
PY
python -c "
import __graft_entry__ as g
g.smoke()"
