#!/bin/bash
mkdir -p gpurun_out
for x in 0 2; do echo "extra=$x"; QA_RELOAD_LAG_EXTRA=$x timeout 300 python scripts/quant_time.py 2>&1 | grep -E "C4|s128k|d256"; done | tee gpurun_out/r02q2_quant.txt
timeout 600 python -m pytest tests/test_quantize_gpu.py -m gpu -q -x -k "long_heads or persistent or garbage" 2>&1 | tail -2
