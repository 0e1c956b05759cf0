"""Developer tool: time the ways of laying the gathered K/V blocks end to end per head (parallel._concat_other_blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import parallel
world, rank, B, H, S, D = 8, 3, 1, 24, 9450, 128
kv_all = torch.randint(0, 255, (world, 2, B, H, S, D), dtype=torch.uint8, device="cuda")
others = [r for r in range(world) if r != rank]

def v_current():
    return parallel._concat_other_blocks(kv_all, rank)

def v_cat_i64():
    w = kv_all.view(torch.int64)
    return torch.cat([w[r] for r in others], dim=3).view(torch.uint8)

def v_cat_u8():
    return torch.cat([kv_all[r] for r in others], dim=3)

def v_per_rank_c128():
    dst = torch.empty((2, B, H, (world - 1) * S, D), dtype=torch.uint8, device="cuda")
    d6 = dst.view(torch.complex128).view(2, B, H, world - 1, S, -1)
    w = kv_all.view(torch.complex128)
    for j, r in enumerate(others):
        d6[:, :, :, j].copy_(w[r])
    return dst

ref = v_current()
for name, fn in (("current (2 strided int64 copies)", v_current), ("torch.cat int64", v_cat_i64), ("torch.cat uint8", v_cat_u8),
                 ("per-rank complex128 copies", v_per_rank_c128)):
    out = fn()
    assert torch.equal(out, ref), name
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms * 1e3:.0f} us  {2 * ref.numel() / ms / 1e6:.0f} GB/s (read + write)")
