"""Developer tool: where does the head-wise quantiser's time go?  Back-to-back launches vs one event pair per launch,
with and without programmatic dependent launch (QA_PDL), Q+K of C2 / C3."""
import os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
dev = torch.device("cuda:0")
for name, (H, S, D, n) in {"C2": (24, 4608, 128, 2), "C3": (32, 8192, 128, 2)}.items():
    sets = [[torch.randn((1, H, S, D), device=dev, dtype=torch.bfloat16) for _ in range(n)] for _ in range(4)]
    by = n * H * S * D * 3
    for i in range(4):
        _native.quantize_fp8(sets[i % 4], _native.QA_SCALE_HEAD)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    a.record()
    for i in range(reps):
        _native.quantize_fp8(sets[i % 4], _native.QA_SCALE_HEAD)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    _native.quant_events = []
    for i in range(reps):
        _native.quantize_fp8(sets[i % 4], _native.QA_SCALE_HEAD)
    torch.cuda.synchronize()
    ev, _native.quant_events = _native.quant_events, None
    ms_ev = statistics.median(x.elapsed_time(y) for x, y in ev[2:])
    # a dummy kernel between launches (what a real step has: the attention kernel)
    z = torch.zeros(1 << 20, device=dev)
    a.record()
    for i in range(reps):
        _native.quantize_fp8(sets[i % 4], _native.QA_SCALE_HEAD)
        z.add_(1.0)
    b.record()
    torch.cuda.synchronize()
    ms_sp = a.elapsed_time(b) / reps
    a.record()
    for i in range(reps):
        z.add_(1.0)
    b.record()
    torch.cuda.synchronize()
    ms_z = a.elapsed_time(b) / reps
    two = []
    _native.quant_events = []
    for i in range(10):
        _native.quantize_fp8(sets[i % 4], _native.QA_SCALE_HEAD_TWO_PASS)
    torch.cuda.synchronize()
    ev, _native.quant_events = _native.quant_events, None
    ms_2p = statistics.median(x.elapsed_time(y) for x, y in ev[2:])
    print(f"{name} PDL={os.environ.get('QA_PDL', '1')} coop={os.environ.get('QA_RING_COOP', '-')}: back-to-back {ms * 1e3:6.1f} us ({by / ms / 1e6:5.0f} GB/s) | "
          f"event pair per launch {ms_ev * 1e3:6.1f} us | with a small kernel in between {(ms_sp - ms_z) * 1e3:6.1f} us (spacer {ms_z * 1e3:.1f}) | "
          f"two-pass (events) {ms_2p * 1e3:6.1f} us", flush=True)
    del sets
