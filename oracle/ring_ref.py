"""CPU restatements used to check the sequence ring (TEST INFRASTRUCTURE ONLY).

The reference has no sequence-parallel path; what is pinned here is the algebra the ring relies on:
  * quantising with a given head scale is the second half of src/quantum_attn/nn.py:14-19 (``t / scale`` clamped
    to +-448, cast to e4m3fn),
  * attention over a key block with its log-sum-exp, and the exact log-sum-exp combine of two blocks - which together
    must reproduce ``fp8_attention_ref`` (src/quantum_attn/ops.py:64-95) on the concatenated keys.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .e4m3 import E4M3_MAX, e4m3_encode_rne_sat
from .quantize_ref import dequantize


def quantize_with_scale(x: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """x [B,H,S,D] float, scale [B,H] fp32 -> e4m3 bytes (fp32 IEEE division, clamp, RNE)."""
    x = np.asarray(x).astype(np.float32)
    s = np.asarray(scale, dtype=np.float32)[..., None, None]
    y = np.clip((x / s).astype(np.float32), -E4M3_MAX, E4M3_MAX)
    return e4m3_encode_rne_sat(y)


def head_scales(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x).astype(np.float32)
    amax = np.abs(x).max(axis=(-2, -1))
    return np.maximum((amax * np.float32(1.0 / E4M3_MAX)).astype(np.float32), np.float32(np.finfo(np.float32).eps))


def attention_block_ref(q8, k8, v8, sq, sk, sv, *, sm_scale=None):
    """Non-causal attention of dequantised q against one key/value block -> (O fp64 [B,H,Sq,D], LSE fp64 [B,H,Sq])."""
    q = torch.from_numpy(dequantize(np.asarray(q8), np.asarray(sq))).double()
    k = torch.from_numpy(dequantize(np.asarray(k8), np.asarray(sk))).double()
    v = torch.from_numpy(dequantize(np.asarray(v8), np.asarray(sv))).double()
    sm = (1.0 / math.sqrt(q.shape[-1])) if sm_scale is None else float(sm_scale)
    s = (q @ k.transpose(-1, -2)) * sm
    lse = torch.logsumexp(s, dim=-1)
    return torch.softmax(s, dim=-1) @ v, lse


def merge_ref(o_a, lse_a, o_b, lse_b):
    """Exact combine of two partial results over disjoint key sets (torch fp64)."""
    m = torch.maximum(lse_a, lse_b)
    wa, wb = torch.exp(lse_a - m), torch.exp(lse_b - m)
    wa = torch.where(torch.isinf(lse_a) & (lse_a < 0), torch.zeros_like(wa), wa)
    wb = torch.where(torch.isinf(lse_b) & (lse_b < 0), torch.zeros_like(wb), wb)
    ws = wa + wb
    safe = torch.where(ws > 0, ws, torch.ones_like(ws))
    o = (wa[..., None] * o_a + wb[..., None] * o_b) / safe[..., None]
    lse = torch.where(ws > 0, m + torch.log(safe), torch.full_like(m, float("-inf")))
    return o, lse
