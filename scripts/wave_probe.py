"""Developer probe: the sequence-sharded call attends head groups in separate launches of exactly one (or two) waves.
How much does that cost against ONE launch of the same CTAs (dynamic dispatch lets fast SMs take more CTAs)?
One rank's share of C4 at N = 8: Sq = 9450 rows per head against all 75600 keys, 24 heads, default mode."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
dev = torch.device("cuda:0")
H, Sq, Skv, D = 24, int(os.environ.get("SQ", 9450)), 75600, 128
q = torch.randn((1, H, Sq, D), device=dev, dtype=torch.bfloat16)
k = torch.randn((1, H, Skv, D), device=dev, dtype=torch.bfloat16)
v = torch.randn((1, H, Skv, D), device=dev, dtype=torch.bfloat16)
(q8,), (sq,) = _native.quantize_fp8([q], _native.QA_SCALE_HEAD)
(k8,), (sk,) = _native.quantize_fp8([k], _native.QA_SCALE_HEAD)
out = torch.empty_like(q)
def run(groups):
    hs = H // groups
    for g in range(groups):
        lo, hi = g * hs, (g + 1) * hs
        _native.fp8_attn_fwd(q8[:, lo:hi], k8[:, lo:hi], v[:, lo:hi], sq[:, lo:hi], sk[:, lo:hi], None, scale_mode=0, is_causal=False,
                             sm_scale=1 / math.sqrt(D), p_mode=2, out_dtype=torch.bfloat16, out=out[:, lo:hi])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}
for rep in range(3):
    for groups in (1, 3, 6, 12, 24):
        run(groups)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(4):
            run(groups)
        e1.record()
        torch.cuda.synchronize()
        res.setdefault(groups, []).append(e0.elapsed_time(e1) / 4)
for g, ts in res.items():
    print(f"Sq={Sq} {g:2d} launches of {H // g * ((Sq + 255) // 256):4d} CTAs: " + " ".join(f"{t:.3f}" for t in ts) + " ms")
