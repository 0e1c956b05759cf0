#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -4 gpurun_out/r02r_pytest.log
for c in 0 1 0 1; do QA_RING_COOP=$c timeout 300 python scripts/quant_time.py 2>&1 | grep -E "^C2|^C3" | sed "s/$/ coop=$c/"; done | tee gpurun_out/r02r_coop.txt
