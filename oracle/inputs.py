"""Seeded synthetic inputs shared by tests, smoke() and bench.py: generated on the CPU generator
so the CPU oracle and the GPU path see identical bits."""
from __future__ import annotations

import torch

# BASELINE.json configs: name -> (B, H, S, D, causal)
CONFIGS = {
    "C1": (2, 8, 512, 64, True),
    "C2_flux": (1, 24, 4608, 128, False),
    "C3_llama": (1, 32, 8192, 128, True),
    "C4_video": (1, 24, 75600, 128, False),
}


def make_qkv(B, H, Sq, Skv, D, *, dtype=torch.bfloat16, seed=0, kind="randn"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(B, H, Sq, D, generator=g, dtype=torch.float32)
    k = torch.randn(B, H, Skv, D, generator=g, dtype=torch.float32)
    v = torch.randn(B, H, Skv, D, generator=g, dtype=torch.float32)
    if kind == "outlier_channels":  # DiT-like heavy-tailed channels
        ch = torch.exp(2.0 * torch.randn(1, 1, 1, D, generator=g))
        q, k, v = q * ch, k * ch, v * ch
    elif kind == "huge_token":  # one large-magnitude token per head: exercises clamp / saturation
        q[:, :, Sq // 3, :] *= 300.0
        k[:, :, Skv // 2, :] *= 300.0
        v[:, :, Skv // 5, :] *= 300.0
    elif kind == "zero_head":  # an all-zero head: scale clamps to eps
        q[:, 0] = 0
        k[:, 0] = 0
        v[:, 0] = 0
    elif kind != "randn":
        raise ValueError(kind)
    return q.to(dtype), k.to(dtype), v.to(dtype)
