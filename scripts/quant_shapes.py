"""Developer tool: quantiser time (memset + kernel, Q/K/V in one call) over head sizes at a fixed total size."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
dev = torch.device("cuda:0")
D = int(os.environ.get("QD", 128))
TOK = 1 << 21
for S in (256, 512, 1024, 2048, 4096, 8192, 32768, 65536):
    BH = TOK // S
    q, k, v = (torch.randn((1, BH, S, D), device=dev, dtype=torch.bfloat16) for _ in range(3))
    row = []
    for mode in (_native.QA_SCALE_HEAD, _native.QA_SCALE_HEAD_TWO_PASS):
        _native.quant_events = []
        for _ in range(6):
            _native.quantize_fp8([q, k, v], mode)
        torch.cuda.synchronize()
        ev, _native.quant_events = _native.quant_events, None
        ms = statistics.median(a.elapsed_time(b) for a, b in ev[1:])
        row.append(f"{ms * 1e3:8.1f} us {3 * BH * S * D * 3 / ms / 1e6:6.0f} GB/s (launches {_native.last_launch_count()})")
    print(f"D={D} S={S:6d} BH={BH:5d} | single-pass {row[0]} | two-pass {row[1]}", flush=True)
