#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ring_gpu.py -m gpu -q -x > gpurun_out/r02_gate_pytest.log 2>&1; tail -5 gpurun_out/r02_gate_pytest.log
