"""Developer tool: head-wise quantiser, mean launch duration over back-to-back launches (Q and K of C2 / C3 / a C4 slab)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
dev = torch.device("cuda:0")
for name, (H, S, D, n) in {"C2": (24, 4608, 128, 2), "C3": (32, 8192, 128, 2), "C4": (24, 75600, 128, 2),
                            "C2x3": (24, 4608, 128, 3)}.items():
    sets = [[torch.randn((1, H, S, D), device=dev, dtype=torch.bfloat16) for _ in range(n)] for _ in range(3)]
    for i in range(3):
        _native.quantize_fp8(sets[i % 3], _native.QA_SCALE_HEAD)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    a.record()
    for i in range(reps):
        _native.quantize_fp8(sets[i % 3], _native.QA_SCALE_HEAD)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    by = n * H * S * D * 3
    print(f"{name}: {ms * 1e3:7.1f} us  {by / ms / 1e6:6.0f} GB/s  launches/call {_native.last_launch_count()} "
          f"coop={os.environ.get('QA_RING_COOP', '-')}", flush=True)
    del sets
