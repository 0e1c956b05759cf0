#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_quantize_gpu.py tests/test_graph_compile_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
tail -5 gpurun_out/r02q_pytest.log
for r in 1 0; do QA_QUANT_RELOAD=$r timeout 300 python scripts/quant_time.py 2>&1 | tail -7; done | tee gpurun_out/r02q_quant.txt
QA_NATIVE_LIB=$PWD/quantumattention_b200/libqattn_sm100_synccheck.so timeout 600 compute-sanitizer --tool synccheck --print-limit 10 python scripts/sanitize_small.py > gpurun_out/sanitizer_synccheck_r02b.txt 2>&1; echo "synccheck rc=$?"; grep -E "Barrier error|at qa|ERROR SUMMARY|sanitize_small" gpurun_out/sanitizer_synccheck_r02b.txt | sort | uniq -c | head
