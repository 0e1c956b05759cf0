"""Dispatch, validation and the public quantiser - the host-side mirror of the reference's ``quantum_attn.nn``
(reference: src/quantum_attn/nn.py).  Same function names, argument meaning and error behaviour:

  * ``can_use_attention(...) -> (bool, reason)``            (reference :282-307, validators :52-205)
  * ``fp8_attention(...)`` raises ``ValueError(reason)``     (reference :433-539)
  * ``dynamically_quantize_fp8(t, reduction_dim=-1)``        (reference :14-42)

What differs, on purpose: no torch.compile / Inductor in the path (the reference compiles a wrapper around its op,
:518-539); the device gate is "sm_100 exactly" instead of ">= 9.0" (:214); an explicit softmax ``scale`` and GQA
(Hq a multiple of Hkv) are accepted; 16-bit q/k passed together with scales are rejected (undefined in the reference,
SURVEY.md Appendix B).
"""
from __future__ import annotations

import functools
import math
from typing import Optional, Tuple

import torch

from . import _native, config, ops

_SUPPORTED_HEAD_DIMS = (64, 128, 256)
_SCALING_METHODS = ("head-wise", "token-wise")


# ------------------------------------------------------------------------------------------------ quantiser
def _dynamically_quantize_fp8(t: torch.Tensor, *, reduction_dim=-1) -> Tuple[torch.Tensor, torch.Tensor]:
    """aten statement of the quantiser (fp32 intermediates); used for fake tensors / shape inference only."""
    eps = torch.finfo(torch.float32).eps
    q_max = torch.finfo(torch.float8_e4m3fn).max
    tf = t.float()
    scale = tf.abs().amax(reduction_dim, keepdim=True).mul(1.0 / q_max).clamp_min(eps)
    t8 = (tf / scale).clamp(-q_max, q_max).to(torch.float8_e4m3fn)
    return t8, scale.squeeze(reduction_dim)


def _normalize_dims(reduction_dim, ndim):
    dims = reduction_dim if isinstance(reduction_dim, (list, tuple)) else [reduction_dim]
    return sorted(d % ndim for d in dims)


def dynamically_quantize_fp8(t: torch.Tensor, *, reduction_dim=-1) -> Tuple[torch.Tensor, torch.Tensor]:
    """e4m3 quantisation with dynamic scales; returns ``(t_fp8, scale_fp32)`` with the reduced dims squeezed.

    Runs the sm_100a quantise kernel.  Supported: a CUDA bf16/fp16 tensor whose last dim is 64/128/256, reducing
    over the last dim (token-wise) or the last two dims (head-wise).  Anything else raises ``ValueError``.
    """
    from torch._subclasses.fake_tensor import is_fake

    if is_fake(t) or torch.compiler.is_dynamo_compiling():
        return _dynamically_quantize_fp8(t, reduction_dim=reduction_dim)
    dims = _normalize_dims(reduction_dim, t.dim())
    if t.device.type != "cuda":
        raise ValueError("Expected the tensor to be on a CUDA device")
    if t.dtype not in (torch.float16, torch.bfloat16):
        raise ValueError(f"Expected dtype torch.float16 or torch.bfloat16, but got {t.dtype} instead.")
    if t.dim() < 2 or t.size(-1) not in _SUPPORTED_HEAD_DIMS:
        raise ValueError(f"Unsupported head dimension: {t.size(-1) if t.dim() else None}")
    if dims == [t.dim() - 1]:
        mode, lead = _native.QA_SCALE_TOKEN, t.shape[:-1]
        x = t.reshape(1, 1, -1, t.size(-1))
    elif dims == [t.dim() - 2, t.dim() - 1]:
        mode, lead = _native.QA_SCALE_HEAD, t.shape[:-2]
        x = t.reshape(1, -1, t.size(-2), t.size(-1)) if t.dim() != 4 else t
    else:
        raise ValueError(f"Unsupported reduction_dim: {reduction_dim}")
    (x8,), (scale,) = _native.quantize_fp8([x], mode)
    return x8.reshape(t.shape), scale.reshape(lead)


# ------------------------------------------------------------------------------------------------ validation
@functools.lru_cache(maxsize=None)
def _device_supported(index: int) -> Tuple[bool, str]:
    major, minor = torch.cuda.get_device_capability(index)
    if major != 10:
        return False, f"CUDA capability 10.x (B200, sm_100) is required, got {major}.{minor}"
    return True, ""


def _pre_check(device: torch.device) -> Tuple[bool, str]:
    if device.type != "cuda":
        return False, f"Expected device to be on a CUDA device, but got device: {device} instead."
    return _device_supported(device.index if device.index is not None else torch.cuda.current_device())


def _validate_input(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None,
                    scaling_method=None, scale_q=None, scale_k=None) -> Tuple[bool, str]:
    if any(t.requires_grad for t in (query, key, value)):
        return False, "NYI: query, key, and value must be leaf tensors"
    if attn_mask is not None:
        return False, "NYI: attn_mask must be None"
    if dropout_p != 0.0:
        return False, "NYI: dropout_p must be 0.0"
    if scale is not None and not (scale > 0.0):
        return False, "scale must be positive"
    f16 = (torch.float16, torch.bfloat16)
    if scaling_method is None:
        if query.dtype != key.dtype or query.dtype != value.dtype:
            return False, (
                "Expected query, key, and value to have the same dtype, but got "
                f"query.dtype: {query.dtype}, key.dtype: {key.dtype}, and value.dtype: {value.dtype} instead."
            )
        if query.dtype not in f16:
            return False, (
                "Expected query, key, and value to have dtype torch.float16 or torch.bfloat16, "
                f"but got query.dtype: {query.dtype} instead."
            )
    else:
        if scaling_method not in _SCALING_METHODS:
            return False, f"Unsupported scaling_method: {scaling_method}"
        if query.dtype not in f16 + (torch.float8_e4m3fn,):
            return False, (
                "Expected query to have dtype torch.float16, torch.bfloat16, or torch.float8_e4m3fn, "
                f"but got query.dtype: {query.dtype} instead."
            )
        if (query.dtype == torch.float8_e4m3fn) != (scale_q is not None and scale_k is not None):
            return False, "float8_e4m3fn query/key need scale_q and scale_k, and 16-bit query/key must not pass them"
    if query.dtype != key.dtype:
        return False, (
            "Expected query and key to have the same dtype, but got "
            f"query.dtype: {query.dtype}, key.dtype: {key.dtype} instead."
        )
    if value.dtype not in f16:
        return False, (
            f"Expected value to have dtype torch.float16 or torch.bfloat16, but got value.dtype: {value.dtype} instead."
        )
    if query.device != key.device or query.device != value.device:
        return False, (
            "Expected query, key, and value to have the same device type, but got "
            f"query.device: {query.device}, key.device: {key.device}, and value.device: {value.device} instead."
        )
    if query.device.type != "cuda":
        return False, "Expected query, key, and value to be on a CUDA device"
    if query.dim() != 4 or key.dim() != 4 or value.dim() != 4:
        return False, "NYI: query, key, and value must be 4D tensors"
    if key.size(-2) != value.size(-2):
        return False, (
            "Expect key and value to have the same sequence length "
            f"but got Sk={key.size(-2)} and Sv={value.size(-2)}."
        )
    if value.size(-1) != query.size(-1) or key.size(-1) != query.size(-1):
        return False, "NYI: query, key and value must have the same embedding dimension"
    if query.size(0) != key.size(0) or key.size(-3) != value.size(-3) or key.size(0) != value.size(0):
        return False, "Expect query, key and value to agree on batch size, and key/value on the number of heads."
    if query.size(-3) % key.size(-3) != 0:
        return False, (
            "Expect the number of query heads to be a multiple of key/value heads "
            f"but got Hq={query.size(-3)} and Hkv={key.size(-3)}."
        )
    if query.size(-1) not in _SUPPORTED_HEAD_DIMS:
        return False, f"Unsupported head dimension: {query.size(-1)}"
    if query.size(-2) < 1 or key.size(-2) < 1:
        return False, "Empty sequences are not supported"
    return True, ""


def can_use_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None,
                      scaling_method=None, scale_q=None, scale_k=None) -> Tuple[bool, str]:
    if config.attention.skip_supported_check:
        return True, ""
    ok, reason = _pre_check(query.device)
    if ok:
        ok, reason = _validate_input(query, key, value, attn_mask, dropout_p, is_causal, scale, scaling_method,
                                     scale_q, scale_k)
    return (True, "") if ok else (False, f"[sm100_tcgen05: {reason}]")


# ------------------------------------------------------------------------------------------------ entry points
def _fp8_attention_wrapper(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None,
                           scale_q=None, scale_k=None, scaling_method=None):
    """Quantise q and k unless scales came in, then call the op (reference: src/quantum_attn/nn.py:394-430)."""
    if (scale_q is None) != (scale_k is None):
        raise ValueError("scale_q and scale_k must be both provided or both not provided")
    if scale_q is None:
        if scaling_method not in _SCALING_METHODS:
            raise ValueError(f"Unsupported scaling_method: {scaling_method}")
        from torch._subclasses.fake_tensor import is_fake

        if is_fake(query) or torch.compiler.is_dynamo_compiling():
            dims = [query.dim() - 2, query.dim() - 1] if scaling_method == "head-wise" else query.dim() - 1
            query, scale_q = _dynamically_quantize_fp8(query, reduction_dim=dims)
            key, scale_k = _dynamically_quantize_fp8(key, reduction_dim=dims)
        else:
            mode = _native.QA_SCALE_HEAD if scaling_method == "head-wise" else _native.QA_SCALE_TOKEN
            if query.shape[1] == key.shape[1]:
                (query, key), (scale_q, scale_k) = _native.quantize_fp8([query, key], mode)  # one launch pair
            else:
                (query,), (scale_q,) = _native.quantize_fp8([query], mode)
                (key,), (scale_k,) = _native.quantize_fp8([key], mode)
    return torch.ops.quantum_attn.fp8_attention_forward(
        query, key, value, scale_q, scale_k, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
        scale=scale,
    )


def _fp8_attention_direct(query, key, value, is_causal, scale, scale_q, scale_k, scaling_method):
    """Eager hot path: no dispatcher hop - quantise everything that needs it in ONE launch pair, then the kernel.

    Launch sequence for 16-bit q, k, v in "fp8" mode: memset, amax(q,k,v), quantise(q,k,v), attention.
    """
    mode = _native.QA_SCALE_HEAD if scaling_method == "head-wise" else _native.QA_SCALE_TOKEN
    p_mode = ops.pv_mode_code()
    v_fp8 = p_mode != _native.QA_P_16BIT
    scale_v = None
    out_dtype = value.dtype if value.dtype in (torch.float16, torch.bfloat16) else torch.bfloat16
    if scale_q is None:
        same_heads = query.shape[1] == key.shape[1]
        if v_fp8 and same_heads and mode == _native.QA_SCALE_HEAD and value.dtype == query.dtype:
            (query, key, value), (scale_q, scale_k, scale_v) = _native.quantize_fp8([query, key, value], mode)
        elif same_heads:
            (query, key), (scale_q, scale_k) = _native.quantize_fp8([query, key], mode)
        else:
            (query,), (scale_q,) = _native.quantize_fp8([query], mode)
            (key,), (scale_k,) = _native.quantize_fp8([key], mode)
    if v_fp8 and scale_v is None:
        (value,), (scale_v,) = _native.quantize_fp8([value], _native.QA_SCALE_HEAD)

    sm_scale = (1.0 / math.sqrt(query.size(-1))) if scale is None else float(scale)
    return _native.fp8_attn_fwd(query, key, value, scale_q, scale_k, scale_v, scale_mode=mode, is_causal=is_causal,
                                sm_scale=sm_scale, p_mode=p_mode, out_dtype=out_dtype)


def fp8_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None, scale_q=None,
                  scale_k=None, scaling_method=None) -> torch.Tensor:
    supported, reason = can_use_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
        scaling_method=scaling_method, scale_q=scale_q, scale_k=scale_k,
    )
    if not supported:
        raise ValueError(reason)
    from torch._subclasses.fake_tensor import is_fake

    traced = torch.compiler.is_dynamo_compiling() or any(
        is_fake(x) for x in (query, key, value, scale_q, scale_k) if x is not None)
    if not traced and not config.attention.force_eager_fallback:
        if (scale_q is None) != (scale_k is None):
            raise ValueError("scale_q and scale_k must be both provided or both not provided")
        return _fp8_attention_direct(query, key, value, is_causal, scale, scale_q, scale_k, scaling_method)
    return _fp8_attention_wrapper(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
        scale_q=scale_q, scale_k=scale_k, scaling_method=scaling_method,
    )


def attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None) -> torch.Tensor:
    """16-bit attention entry point (reference: src/quantum_attn/nn.py:325-391): bf16 / fp16 q, k, v straight into the
    fused kernel (kind::f16 MMAs), no quantisation."""
    supported, reason = can_use_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale
    )
    if not supported:
        raise ValueError(f"Unsupported input: {reason}")
    from torch._subclasses.fake_tensor import is_fake

    traced = torch.compiler.is_dynamo_compiling() or any(is_fake(x) for x in (query, key, value))
    if not traced and not config.attention.force_eager_fallback:
        return ops.attention_native(query, key, value, is_causal=is_causal, scale=scale)  # no dispatcher hop
    return torch.ops.quantum_attn.attention_forward(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale
    )
