"""fp32 restatement of the reference's dynamic FP8 quantiser.

Follows src/quantum_attn/nn.py:14-19 (``_dynamically_quantize_fp8``) with the reduction dims the attention wrapper
uses (src/quantum_attn/nn.py:410-418): "head-wise" reduces over (S, D), "token-wise" over D.  "block-128" (one scale
per 128 consecutive tokens) is this repo's extension for K/V tiles and follows the same formula.

    scale = max(amax(|t|) * fp32(1/448), fp32_eps)        # fp32 multiply, then clamp_min
    t8    = e4m3_rne(clamp(t / scale, -448, 448))          # IEEE fp32 division

All intermediates are fp32 - the arithmetic the reference's default (Inductor-fused) path performs; its eager path
keeps bf16 intermediates and differs on ~2-3 % of bytes, so byte parity is defined against
this restatement and pinned by tests/golden/quantize_*.npz (generated from the reference run on fp32 inputs).
"""
from __future__ import annotations

import numpy as np

from .e4m3 import E4M3_MAX, e4m3_decode, e4m3_encode_rne_sat

_EPS = np.float32(np.finfo(np.float32).eps)
_INV_QMAX = np.float32(1.0 / E4M3_MAX)


def _reduce_axes(ndim: int, mode: str):
    if mode == "head-wise":
        return (ndim - 2, ndim - 1)
    if mode == "token-wise":
        return (ndim - 1,)
    raise ValueError(f"Unsupported scaling_method: {mode}")


def quantize_fp8(x: np.ndarray, mode: str = "head-wise", block: int = 128):
    """x: float array [..., S, D] (any float dtype; widened to fp32 first).

    Returns (bytes uint8 same shape, scale float32 with the reduced dims squeezed).
    """
    x = np.asarray(x).astype(np.float32)
    if mode == "block-128":
        *lead, S, D = x.shape
        nb = (S + block - 1) // block
        pad = nb * block - S
        xp = np.pad(x, [(0, 0)] * len(lead) + [(0, pad), (0, 0)])
        xb = xp.reshape(*lead, nb, block, D)
        amax = np.abs(xb).max(axis=(-2, -1), keepdims=True)
        scale = np.maximum(amax * _INV_QMAX, _EPS).astype(np.float32)
        y = np.clip((xb / scale).astype(np.float32), -E4M3_MAX, E4M3_MAX)
        b = e4m3_encode_rne_sat(y).reshape(*lead, nb * block, D)[..., :S, :]
        return b, scale.reshape(*lead, nb)
    axes = _reduce_axes(x.ndim, mode)
    amax = np.abs(x).max(axis=axes, keepdims=True)
    scale = np.maximum((amax * _INV_QMAX).astype(np.float32), _EPS).astype(np.float32)
    y = np.clip((x / scale).astype(np.float32), -E4M3_MAX, E4M3_MAX)
    return e4m3_encode_rne_sat(y), np.squeeze(scale, axis=axes)


def dequantize(b: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """bytes [..., S, D] and scale [...] or [..., S] -> fp32."""
    v = e4m3_decode(b)
    s = np.asarray(scale, dtype=np.float32)
    while s.ndim < v.ndim:
        s = s[..., None]
    return (v * s).astype(np.float32)
