"""fp32/fp64 restatement of the reference's FP8 attention semantics.

``fp8_attention_ref`` follows src/quantum_attn/ops.py:64-95 (``_fp8_attention_forward``): dequantise q and k with
their scales, then scaled-dot-product attention with the value tensor, top-left aligned causal mask
(src/quantum_attn/tk/attention.py:252-263), softmax scale 1/sqrt(D) unless ``scale`` is given
(src/quantum_attn/tk/attention.py:208-210).  The reference performs the dequantisation and SDPA in v.dtype (bf16);
this restatement keeps everything in fp32 (fp64 accumulation optional), i.e. it is the exact-arithmetic version of the
same definition, which is what BASELINE.json names as the comparison target ("fp32 SDPA on the dequantised inputs").
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .quantize_ref import dequantize


def sdpa_ref(q, k, v, *, is_causal=False, scale=None, dtype=torch.float64, chunk=1024):
    """q [B,H,Sq,D], k/v [B,H,Skv,D] (torch or numpy) -> torch tensor [B,H,Sq,D] in ``dtype``.

    Chunked over query rows so large S fits in memory; row-wise stable softmax.
    """
    q = torch.as_tensor(np.asarray(q) if not torch.is_tensor(q) else q).to(dtype)
    k = torch.as_tensor(np.asarray(k) if not torch.is_tensor(k) else k).to(dtype)
    v = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(dtype)
    B, H, Sq, D = q.shape
    Skv = k.shape[2]
    sm = (1.0 / math.sqrt(D)) if scale is None else float(scale)
    out = torch.empty(B, H, Sq, v.shape[-1], dtype=dtype)
    kt = k.transpose(-1, -2)
    cols = torch.arange(Skv)
    for r0 in range(0, Sq, chunk):
        r1 = min(Sq, r0 + chunk)
        s = (q[:, :, r0:r1] @ kt) * sm
        if is_causal:
            rows = torch.arange(r0, r1)
            mask = cols[None, :] > rows[:, None]
            s = s.masked_fill(mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[:, :, r0:r1] = p @ v
    return out


def fp8_attention_ref(q8, k8, v, scale_q, scale_k, *, scale_v=None, is_causal=False, scale=None,
                      dtype=torch.float64):
    """q8/k8: uint8 e4m3 bytes [B,H,S,D]; scale_q/scale_k: fp32 [B,H] (head-wise) or [B,H,S] (token-wise).

    v: float array (unquantised, the reference's semantics) or uint8 e4m3 bytes with ``scale_v`` (this repo's
    FP8-V mode; dequantised first).
    """
    qh = dequantize(np.asarray(q8), np.asarray(scale_q))
    kh = dequantize(np.asarray(k8), np.asarray(scale_k))
    if scale_v is not None:
        vh = dequantize(np.asarray(v), np.asarray(scale_v))
    else:
        vh = v.float().numpy() if torch.is_tensor(v) else np.asarray(v, dtype=np.float32)
    return sdpa_ref(qh, kh, vh, is_causal=is_causal, scale=scale, dtype=dtype)


def attention_flops(B, H, Sq, Skv, D, causal):
    """Algorithmic FLOPs, the reference's own formula (tests/test_interface.py:121-125)."""
    f = 4 * B * H * Sq * Skv * D
    return f // 2 if causal else f


def cpu_reference_step(q8, k8, v, scale_q, scale_k, *, is_causal=False, scale=None):
    """The reference's op definition executed on the host cores (the CPU arm of bench.py).

    Follows src/quantum_attn/ops.py:64-95 literally - cast q8/k8 up, multiply by the scales, aten SDPA - but in fp32
    and with the MATH backend, which is the arithmetic BASELINE.json names for the CPU baseline.  Inputs are torch
    CPU tensors: q8/k8 float8_e4m3fn, v fp32, scales fp32 ([B,H]).
    """
    from torch.nn.attention import SDPBackend, sdpa_kernel

    q = q8.to(torch.float32) * scale_q[..., None, None]
    k = k8.to(torch.float32) * scale_k[..., None, None]
    with sdpa_kernel(SDPBackend.MATH):
        return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=is_causal, scale=scale)
