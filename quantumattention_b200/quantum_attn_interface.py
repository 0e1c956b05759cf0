"""The seven public entry points, signature-compatible with the reference
(reference: src/quantum_attn/quantum_attn_interface.py:41-248).

``*_with_fallback`` are CompositeImplicitAutograd ops that run the same support check and, when the input is not
supported, call ``F.scaled_dot_product_attention`` on the ORIGINAL tensors - the reference's documented behaviour
(:76-98).  The non-fallback functions raise ``ValueError(reason)`` instead.  (The reference's
``fp8_token_wise_attn_func_with_fallback`` passes a stray ``scaling_method=`` kwarg and raises TypeError on its
supported branch, :229-238; that bug is not reproduced.)
"""
from typing import Optional

import torch
import torch.nn.functional as F

from .nn import QuantizedKV, attention, can_use_attention, dynamically_quantize_fp8, fp8_attention, quantize_kv  # noqa: F401

__all__ = [
    "attn_func",
    "attn_func_with_fallback",
    "fp8_attn_func",
    "fp8_attn_func_with_fallback",
    "fp8_token_wise_attn_func",
    "fp8_token_wise_attn_func_with_fallback",
    "dynamically_quantize_fp8",
]

_SDPA_SIG = ("(Tensor query, Tensor key, Tensor value, Tensor? attn_mask=None, float dropout_p=0.0, "
             "bool is_causal=False, *, float? scale=None{extra}) -> Tensor")


def _composite_op(name: str, extra: str = ""):
    def register(fn):
        torch.library.define(f"quantum_attn::{name}", _SDPA_SIG.format(extra=extra))
        torch.library.impl(f"quantum_attn::{name}", ["CompositeImplicitAutograd"])(fn)
        return getattr(torch.ops.quantum_attn, name)

    return register


def attn_func(query, key, value, attn_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0,
              is_causal: bool = False, *, scale: float = None) -> torch.Tensor:
    return attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale)


@_composite_op("attn_func_with_fallback")
def attn_func_with_fallback(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None):
    if can_use_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                         scale=scale)[0]:
        return attn_func(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                         scale=scale)
    return F.scaled_dot_product_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p,
                                          is_causal=is_causal, scale=scale)


def fp8_attn_func(query, key, value, attn_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0,
                  is_causal: bool = False, *, scale: float = None, scale_q: Optional[torch.Tensor] = None,
                  scale_k: Optional[torch.Tensor] = None, scaling_method: Optional[str] = None,
                  scale_v: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's signature (src/quantum_attn/quantum_attn_interface.py:101-113) plus one trailing optional
    keyword: ``scale_v`` for a value tensor that was quantised ahead of time (``quantize_kv``; FP8 P modes).  The key
    alone may come pre-quantised too (e4m3 ``key`` + ``scale_k`` with a 16-bit ``query``)."""
    return fp8_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
        scale_q=scale_q, scale_k=scale_k, scaling_method="head-wise" if scaling_method is None else scaling_method,
        scale_v=scale_v,
    )


@_composite_op("fp8_attn_func_with_fallback", ", str? scaling_method=None")
def fp8_attn_func_with_fallback(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None,
                                scaling_method=None):
    method = "head-wise" if scaling_method is None else scaling_method
    if can_use_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                         scale=scale, scaling_method=method)[0]:
        return fp8_attn_func(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                             scale=scale, scaling_method=method)
    return F.scaled_dot_product_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p,
                                          is_causal=is_causal, scale=scale)


def fp8_token_wise_attn_func(query, key, value, attn_mask: Optional[torch.Tensor] = None, dropout_p: float = 0.0,
                             is_causal: bool = False, *, scale: float = None,
                             scale_q: Optional[torch.Tensor] = None,
                             scale_k: Optional[torch.Tensor] = None,
                             scale_v: Optional[torch.Tensor] = None) -> torch.Tensor:
    return fp8_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
        scale_q=scale_q, scale_k=scale_k, scaling_method="token-wise", scale_v=scale_v,
    )


@_composite_op("fp8_token_wise_attn_func_with_fallback")
def fp8_token_wise_attn_func_with_fallback(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *,
                                           scale=None):
    if can_use_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                         scale=scale, scaling_method="token-wise")[0]:
        return fp8_token_wise_attn_func(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p,
                                        is_causal=is_causal, scale=scale)
    return F.scaled_dot_product_attention(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p,
                                          is_causal=is_causal, scale=scale)
