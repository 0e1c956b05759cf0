#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -q -k "token or golden or stress or gqa" --maxfail 20 > gpurun_out/r02_tok_pytest.log 2>&1; tail -3 gpurun_out/r02_tok_pytest.log
AB_SHAPES=C2,C2u,C3,d64,d256 AB_MODES=16bit,token python scripts/ab_kernels.py token3 2>&1 | tee gpurun_out/r02_token_ab.txt
