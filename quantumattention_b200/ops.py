"""torch custom ops of the hot path, same names and schemas as the reference so traced graphs look the same
(reference: src/quantum_attn/ops.py:32-147):

    quantum_attn::fp8_attention_forward(query, key, value, scale_q, scale_k, attn_mask, dropout_p, is_causal, *, scale)
    quantum_attn::attention_forward(query, key, value, attn_mask, dropout_p, is_causal, *, scale)

The CUDA implementation calls the sm_100a kernels through the C ABI (``_native``); the reference's CUDA implementation
is aten SDPA on the dequantised inputs, which survives here only as ``fp8_attention_definition`` - the semantic
definition used for fake-tensor shape inference and as an explicit debug comparator.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _native, config

aten = torch.ops.aten

_PV_MODES = {"fp8": _native.QA_P_E4M3, "fp8_hilo": _native.QA_P_E4M3_HILO, "16bit": _native.QA_P_16BIT}


def pv_mode_code(name: Optional[str] = None) -> int:
    name = config.attention.pv_mode if name is None else name
    try:
        return _PV_MODES[name]
    except KeyError:
        raise ValueError(f"Unsupported pv_mode: {name} (expected one of {sorted(_PV_MODES)})") from None


def fp8_attention_definition(query, key, value, scale_q=None, scale_k=None, attn_mask=None, dropout_p=0.0,
                             is_causal=False, *, scale=None):
    """What the op computes, stated with aten ops (reference: src/quantum_attn/ops.py:64-95)."""
    out_dtype = value.dtype
    query = query.to(out_dtype)
    key = key.to(out_dtype)
    if scale_q is not None:
        sq, sk = scale_q.to(out_dtype), scale_k.to(out_dtype)
        while sq.dim() < query.dim():
            sq, sk = sq.unsqueeze(-1), sk.unsqueeze(-1)
        query, key = query * sq, key * sk
    if key.size(1) != query.size(1):  # GQA (extension; the reference's Python gate forbids it)
        rep = query.size(1) // key.size(1)
        key, value = key.repeat_interleave(rep, 1), value.repeat_interleave(rep, 1)
    return aten.scaled_dot_product_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale
    ).contiguous()


def attention_definition(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None):
    """What the 16-bit op computes, stated with aten ops (reference: src/quantum_attn/ops.py:27-29)."""
    if key.size(1) != query.size(1):  # GQA (extension; the reference's Python gate forbids it)
        rep = query.size(1) // key.size(1)
        key, value = key.repeat_interleave(rep, 1), value.repeat_interleave(rep, 1)
    return aten.scaled_dot_product_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale
    ).contiguous()


def attention_native(query, key, value, is_causal=False, scale=None, return_lse=False):
    """16-bit q, k, v -> attention output through the sm_100a kernel (no fallback)."""
    if query.dtype not in (torch.float16, torch.bfloat16) or key.dtype != query.dtype or value.dtype != query.dtype:
        raise ValueError("attention_forward expects query, key and value of one dtype, torch.float16 or torch.bfloat16")
    sm_scale = (1.0 / math.sqrt(query.size(-1))) if scale is None else float(scale)
    return _native.attn_fwd(query, key, value, is_causal=is_causal, sm_scale=sm_scale, return_lse=return_lse)


def _scale_mode_of(scale_q: torch.Tensor, query: torch.Tensor) -> int:
    if scale_q.dim() == query.dim() - 2:
        return _native.QA_SCALE_HEAD
    if scale_q.dim() == query.dim() - 1:
        return _native.QA_SCALE_TOKEN
    raise ValueError(f"scale_q must have rank {query.dim() - 2} (head-wise) or {query.dim() - 1} (token-wise)")


def fp8_attention_native(query, key, value, scale_q, scale_k, is_causal=False, scale=None, pv_mode=None,
                         return_lse=False):
    """q8/k8 (e4m3) + scales + 16-bit V -> attention output through the sm_100a kernels (no fallback)."""
    if query.dtype != torch.float8_e4m3fn or key.dtype != torch.float8_e4m3fn:
        raise ValueError("fp8_attention_forward expects float8_e4m3fn query and key")
    if scale_q is None or scale_k is None:
        raise ValueError("fp8_attention_forward needs scale_q and scale_k")
    mode = pv_mode_code(pv_mode)
    scale_mode = _scale_mode_of(scale_q, query)
    sm_scale = (1.0 / math.sqrt(query.size(-1))) if scale is None else float(scale)
    if mode == _native.QA_P_16BIT:
        v_in, scale_v = value, None
    elif value.dtype == torch.float8_e4m3fn:
        raise ValueError("an e4m3 value tensor needs its scale; use quantumattention_b200._native.fp8_attn_fwd")
    else:
        (v_in,), (scale_v,) = _native.quantize_fp8([value], _native.QA_SCALE_HEAD)
    out_dtype = value.dtype if value.dtype in (torch.float16, torch.bfloat16) else torch.bfloat16
    return _native.fp8_attn_fwd(query, key, v_in, scale_q, scale_k, scale_v, scale_mode=scale_mode,
                                is_causal=is_causal, sm_scale=sm_scale, p_mode=mode, out_dtype=out_dtype,
                                return_lse=return_lse)


@torch.library.custom_op("quantum_attn::fp8_attention_forward", mutates_args=(), device_types=("cuda",))
def fp8_attention_forward(
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    scale_q: Optional[torch.Tensor] = None,
    scale_k: Optional[torch.Tensor] = None,
    attn_mask: Optional[torch.Tensor] = None,
    dropout_p: float = 0.0,
    is_causal: bool = False,
    *,
    scale: Optional[float] = None,
) -> torch.Tensor:
    if attn_mask is not None or dropout_p != 0.0:
        raise ValueError("NYI: attn_mask must be None and dropout_p must be 0.0")
    if config.attention.force_eager_fallback:
        return fp8_attention_definition(query, key, value, scale_q, scale_k, attn_mask, dropout_p, is_causal,
                                        scale=scale)
    return fp8_attention_native(query, key, value, scale_q, scale_k, is_causal=is_causal, scale=scale)


@torch.library.register_fake("quantum_attn::fp8_attention_forward")
def _(query, key, value, scale_q=None, scale_k=None, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None):
    return torch.empty(query.shape[:-1] + (value.shape[-1],), dtype=value.dtype, device=query.device)


@torch.library.custom_op("quantum_attn::attention_forward", mutates_args=(), device_types=("cuda",))
def attention_forward(
    query: torch.Tensor,
    key: torch.Tensor,
    value: torch.Tensor,
    attn_mask: Optional[torch.Tensor] = None,
    dropout_p: float = 0.0,
    is_causal: bool = False,
    *,
    scale: Optional[float] = None,
) -> torch.Tensor:
    # 16-bit QK^T path (reference: src/quantum_attn/ops.py:32-45): the same fused kernel with kind::f16 MMAs
    if attn_mask is not None or dropout_p != 0.0:
        raise ValueError("NYI: attn_mask must be None and dropout_p must be 0.0")
    if config.attention.force_eager_fallback:
        return attention_definition(query, key, value, attn_mask, dropout_p, is_causal, scale=scale)
    return attention_native(query, key, value, is_causal=is_causal, scale=scale)


@torch.library.register_fake("quantum_attn::attention_forward")
def _(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None):
    return torch.empty(query.shape[:-1] + (value.shape[-1],), dtype=value.dtype, device=query.device)
