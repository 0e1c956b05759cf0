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
from typing import NamedTuple, Optional, Tuple

import torch

from . import _native, config, ops

_SUPPORTED_HEAD_DIMS = (64, 128, 256)
_SCALING_METHODS = ("head-wise", "token-wise")
_F16_OR_E4M3 = (torch.float16, torch.bfloat16, torch.float8_e4m3fn)


def _plain_tensor(x) -> bool:
    """A real tensor for sure (exact type, no functional wrapper): the common eager case, decided without the
    isinstance chain of ``is_fake``."""
    return type(x) is torch.Tensor and not torch._is_functional_tensor(x)


# ------------------------------------------------------------------------------------------------ quantiser
def _dynamically_quantize_fp8(t: torch.Tensor, *, reduction_dim=-1) -> Tuple[torch.Tensor, torch.Tensor]:
    """aten statement of the quantiser (fp32 intermediates); used for fake tensors / shape inference only."""
    eps = torch.finfo(torch.float32).eps
    q_max = torch.finfo(torch.float8_e4m3fn).max
    tf = t.float()
    scale = tf.abs().amax(reduction_dim, keepdim=True).mul(1.0 / q_max).clamp_min(eps)
    t8 = (tf / scale).clamp(-q_max, q_max).to(torch.float8_e4m3fn)
    return t8, scale.squeeze(reduction_dim)


def dynamically_quantize_fp8(t: torch.Tensor, *, reduction_dim=-1) -> Tuple[torch.Tensor, torch.Tensor]:
    """e4m3 quantisation with dynamic scales; returns ``(t_fp8, scale_fp32)`` with the reduced dims squeezed.

    Runs the sm_100a quantise kernel - eagerly, and from a ``torch.compile``d graph too: under tracing the call
    becomes the custom op ``quantum_attn::dynamically_quantize_fp8`` (the reference traces its quantiser into aten ops
    for Inductor, src/quantum_attn/nn.py:22-42).  Supported: a CUDA bf16/fp16 tensor whose last dim is 64/128/256,
    reducing over the last dim (token-wise) or the last two dims (head-wise).  Anything else raises ``ValueError``.
    """
    from torch._subclasses.fake_tensor import is_fake

    if torch.compiler.is_dynamo_compiling() or is_fake(t):
        dims = reduction_dim if isinstance(reduction_dim, (list, tuple)) else [reduction_dim]
        return torch.ops.quantum_attn.dynamically_quantize_fp8(t, [int(d) for d in dims])
    return ops.quantize_native(t, reduction_dim)


class QuantizedKV(NamedTuple):
    """K (and, in the FP8 P modes, V) quantised once, for reuse across attention calls - e.g. the K/V of a cached
    context attended by many query chunks, or by every step of a sampler.  Pass the fields back as
    ``fp8_attn_func(q, kv.key, kv.value, scale_k=kv.scale_k, scale_v=kv.scale_v)``."""
    key: torch.Tensor                 # e4m3 [B, H, S, D]
    scale_k: torch.Tensor             # fp32 [B, H] (head-wise) or [B, H, S] (token-wise)
    value: torch.Tensor               # e4m3 [B, H, S, D] (FP8 P modes) or the 16-bit tensor as it came ("16bit" mode)
    scale_v: Optional[torch.Tensor]   # fp32 [B, H], or None when value is 16-bit


def quantize_kv(key: torch.Tensor, value: torch.Tensor, *, scaling_method: str = "head-wise",
                pv_mode: Optional[str] = None) -> QuantizedKV:
    """Quantise K (and V when the P mode multiplies an e4m3 V) in one launch; see ``QuantizedKV``.  The reference
    quantises K on every call (src/quantum_attn/nn.py:410-418)."""
    if scaling_method not in _SCALING_METHODS:
        raise ValueError(f"Unsupported scaling_method: {scaling_method}")
    mode = _native.QA_SCALE_HEAD if scaling_method == "head-wise" else _native.QA_SCALE_TOKEN
    v_fp8 = ops.pv_mode_code(pv_mode) != _native.QA_P_16BIT
    if v_fp8 and mode == _native.QA_SCALE_HEAD and value.dtype == key.dtype and value.shape == key.shape:
        (k8, v8), (sk, sv) = _native.quantize_fp8([key, value], mode)
        return QuantizedKV(k8, sk, v8, sv)
    (k8,), (sk,) = _native.quantize_fp8([key], mode)
    if not v_fp8:
        return QuantizedKV(k8, sk, value, None)
    (v8,), (sv,) = _native.quantize_fp8([value], _native.QA_SCALE_HEAD)
    return QuantizedKV(k8, sk, v8, sv)


# ------------------------------------------------------------------------------------------------ validation
@torch.compiler.assume_constant_result
def _attention_flag(name: str):
    """``config.attention.<name>`` in a form dynamo can trace through: the config module's sub-config proxies cannot be
    inlined by dynamo (torch/utils/_config_module.py is on its skip list), so under tracing the flag is read once, when
    the graph is built, and baked in as a constant - the same moment the reference's nested torch.compile reads it."""
    return getattr(config.attention, name)


@functools.lru_cache(maxsize=None)
def _device_supported(index: int) -> Tuple[bool, str]:
    major, minor = torch.cuda.get_device_capability(index)
    if major != 10:
        return False, f"CUDA capability 10.x (B200, sm_100) is required, got {major}.{minor}"
    return True, ""


@torch.compiler.assume_constant_result
def _device_supported_traced(index: Optional[int]) -> Tuple[bool, str]:
    # (under dynamo the capability query would otherwise be traced INTO the graph as a call of its own)
    return _device_supported(index if index is not None else torch.cuda.current_device())


def _pre_check(device: torch.device) -> Tuple[bool, str]:
    if device.type != "cuda":
        return False, f"Expected device to be on a CUDA device, but got device: {device} instead."
    if torch.compiler.is_dynamo_compiling():
        return _device_supported_traced(device.index)
    return _device_supported(device.index if device.index is not None else torch.cuda.current_device())


def _validate_input(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None,
                    scaling_method=None, scale_q=None, scale_k=None, scale_v=None) -> Tuple[bool, str]:
    # (the eager hot path runs through here on every call: attributes are fetched once)
    if query.requires_grad or key.requires_grad or value.requires_grad:
        return False, "NYI: query, key, and value must be leaf tensors"
    qdt, kdt, vdt = query.dtype, key.dtype, value.dtype
    if attn_mask is not None:
        return False, "NYI: attn_mask must be None"
    if dropout_p != 0.0:
        return False, "NYI: dropout_p must be 0.0"
    if scale is not None and not (scale > 0.0):
        return False, "scale must be positive"
    f16 = (torch.float16, torch.bfloat16)
    if scaling_method is None:
        if qdt != kdt or qdt != vdt:
            return False, (
                "Expected query, key, and value to have the same dtype, but got "
                f"query.dtype: {query.dtype}, key.dtype: {key.dtype}, and value.dtype: {value.dtype} instead."
            )
        if qdt not in f16:
            return False, (
                "Expected query, key, and value to have dtype torch.float16 or torch.bfloat16, "
                f"but got query.dtype: {query.dtype} instead."
            )
    else:
        if scaling_method not in _SCALING_METHODS:
            return False, f"Unsupported scaling_method: {scaling_method}"
        if qdt not in _F16_OR_E4M3:
            return False, (
                "Expected query to have dtype torch.float16, torch.bfloat16, or torch.float8_e4m3fn, "
                f"but got query.dtype: {query.dtype} instead."
            )
        # a tensor is pre-quantised exactly when its scale comes with it (16-bit q/k WITH scales is undefined in the
        # reference); the key alone may be pre-quantised (K8 reuse: the query is quantised here)
        if (qdt == torch.float8_e4m3fn) != (scale_q is not None) or \
                (kdt == torch.float8_e4m3fn) != (scale_k is not None):
            return False, "float8_e4m3fn query/key need scale_q and scale_k, and 16-bit query/key must not pass them"
        if scale_q is not None and scale_k is None:
            return False, "a pre-quantised query needs a pre-quantised key (scale_q and scale_k)"
        if kdt not in _F16_OR_E4M3:
            return False, (
                "Expected key to have dtype torch.float16, torch.bfloat16, or torch.float8_e4m3fn, "
                f"but got key.dtype: {key.dtype} instead."
            )
        if (vdt == torch.float8_e4m3fn) != (scale_v is not None):
            return False, "a float8_e4m3fn value needs scale_v, and a 16-bit value must not pass it"
    if qdt != kdt and not (scaling_method is not None and kdt == torch.float8_e4m3fn):
        return False, (
            "Expected query and key to have the same dtype, but got "
            f"query.dtype: {query.dtype}, key.dtype: {key.dtype} instead."
        )
    if vdt not in f16 and not (scaling_method is not None and scale_v is not None):
        return False, (
            f"Expected value to have dtype torch.float16 or torch.bfloat16, but got value.dtype: {value.dtype} instead."
        )
    qdev = query.device
    if qdev != key.device or qdev != value.device:
        return False, (
            "Expected query, key, and value to have the same device type, but got "
            f"query.device: {query.device}, key.device: {key.device}, and value.device: {value.device} instead."
        )
    if qdev.type != "cuda":
        return False, "Expected query, key, and value to be on a CUDA device"
    qs, ks, vs = query.shape, key.shape, value.shape
    if len(qs) != 4 or len(ks) != 4 or len(vs) != 4:
        return False, "NYI: query, key, and value must be 4D tensors"
    if ks[2] != vs[2]:
        return False, (
            "Expect key and value to have the same sequence length "
            f"but got Sk={ks[2]} and Sv={vs[2]}."
        )
    if vs[3] != qs[3] or ks[3] != qs[3]:
        return False, "NYI: query, key and value must have the same embedding dimension"
    if qs[0] != ks[0] or ks[1] != vs[1] or ks[0] != vs[0]:
        return False, "Expect query, key and value to agree on batch size, and key/value on the number of heads."
    if qs[1] % ks[1] != 0:
        return False, (
            "Expect the number of query heads to be a multiple of key/value heads "
            f"but got Hq={qs[1]} and Hkv={ks[1]}."
        )
    if qs[3] not in _SUPPORTED_HEAD_DIMS:
        return False, f"Unsupported head dimension: {qs[3]}"
    if qs[2] < 1 or ks[2] < 1:
        return False, "Empty sequences are not supported"
    return True, ""


def can_use_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None,
                      scaling_method=None, scale_q=None, scale_k=None, scale_v=None) -> Tuple[bool, str]:
    if _attention_flag("skip_supported_check"):
        return True, ""
    ok, reason = _pre_check(query.device)
    if ok:
        ok, reason = _validate_input(query, key, value, attn_mask, dropout_p, is_causal, scale, scaling_method,
                                     scale_q, scale_k, scale_v)
    return (True, "") if ok else (False, f"[sm100_tcgen05: {reason}]")


# ------------------------------------------------------------------------------------------------ entry points
def _fp8_attention_wrapper(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None,
                           scale_q=None, scale_k=None, scaling_method=None):
    """Quantise q and k unless scales came in, then call the op (reference: src/quantum_attn/nn.py:394-430).  This is
    the TRACED path (dynamo / fake tensors): the graph holds ``quantum_attn::quantize_qk_fp8`` - the hand-written
    quantiser, one launch for Q and K - followed by the reference's own op ``quantum_attn::fp8_attention_forward``."""
    if scale_q is None:
        if scaling_method not in _SCALING_METHODS:
            raise ValueError(f"Unsupported scaling_method: {scaling_method}")
        token = scaling_method == "token-wise"
        if scale_k is None:
            query, key, scale_q, scale_k = torch.ops.quantum_attn.quantize_qk_fp8(query, key, token)
        else:  # pre-quantised key: quantise the query alone
            dims = [query.dim() - 1] if token else [query.dim() - 2, query.dim() - 1]
            query, scale_q = torch.ops.quantum_attn.dynamically_quantize_fp8(query, dims)
    return torch.ops.quantum_attn.fp8_attention_forward(
        query, key, value, scale_q, scale_k, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
        scale=scale,
    )


def _check_method_against_scales(scaling_method, scale_mode):
    want = _native.QA_SCALE_HEAD if scaling_method == "head-wise" else _native.QA_SCALE_TOKEN
    if scale_mode != want:
        raise ValueError(f"scaling_method={scaling_method!r} contradicts the shape of the scales passed in "
                         f"({'[B,H]: head-wise' if scale_mode == _native.QA_SCALE_HEAD else '[B,H,S]: token-wise'})")


def _fp8_attention_direct(query, key, value, is_causal, scale, scale_q, scale_k, scaling_method, scale_v=None):
    """Eager hot path, no dispatcher hop.  16-bit q, k, v: ONE crossing of the C ABI (quantise whatever the P mode
    needs in one launch, then the fused kernel).  Pre-quantised tensors skip their share of the quantiser."""
    mode = _native.QA_SCALE_HEAD if scaling_method == "head-wise" else _native.QA_SCALE_TOKEN
    p_mode = ops.pv_mode_code()
    sm_scale = (1.0 / math.sqrt(query.size(-1))) if scale is None else float(scale)
    if scale_k is None:
        # nothing pre-quantised (the validators guarantee scale_q is None too, and a 16-bit value without scale_v)
        if _native.attn_events is None and _native.quant_events is None and query.dtype == key.dtype == value.dtype:
            return _native.fp8_attn_func(query, key, value, scale_mode=mode, is_causal=is_causal, sm_scale=sm_scale,
                                         p_mode=p_mode)
        # (bench.py brackets the attention launch alone with CUDA events on some steps: two crossings there)
        query, key, scale_q, scale_k = ops.quantize_qk_native(query, key, mode == _native.QA_SCALE_TOKEN)
    else:
        # the granularity of the scales is what their SHAPE says, as in the reference and the traced path
        # (ops._scale_mode_of); scaling_method must agree with it
        if scale_q is None:
            _check_method_against_scales(scaling_method, ops._scale_mode_of(scale_k, key))
            (query,), (scale_q,) = _native.quantize_fp8([query], mode)
        else:
            _check_method_against_scales(scaling_method, ops._scale_mode_of(scale_q, query, scale_k, key))
    out_dtype = value.dtype if value.dtype in (torch.float16, torch.bfloat16) else (
        query.dtype if query.dtype in (torch.float16, torch.bfloat16) else torch.bfloat16)
    return ops.fp8_attention_native(query, key, value, scale_q, scale_k, is_causal=is_causal, scale=scale,
                                    scale_v=scale_v, out_dtype=out_dtype)


def fp8_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, *, scale=None, scale_q=None,
                  scale_k=None, scaling_method=None, scale_v=None) -> torch.Tensor:
    """(reference: src/quantum_attn/nn.py:433-539.)  Extensions over the reference's signature, all optional: the key
    alone may come pre-quantised (e4m3 ``key`` + ``scale_k`` with a 16-bit ``query``), and in the FP8 P modes the
    value too (e4m3 ``value`` + head-wise ``scale_v``) - see ``quantize_kv``."""
    supported, reason = can_use_attention(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale,
        scaling_method=scaling_method, scale_q=scale_q, scale_k=scale_k, scale_v=scale_v,
    )
    if not supported:
        raise ValueError(reason)
    from torch._subclasses.fake_tensor import is_fake

    if torch.compiler.is_dynamo_compiling():
        traced = True
    elif _plain_tensor(query) and _plain_tensor(key) and _plain_tensor(value) and scale_q is None and scale_k is None:
        traced = False
    else:
        traced = any(is_fake(x) for x in (query, key, value, scale_q, scale_k) if x is not None)
    if not traced and not config.attention.force_eager_fallback:
        return _fp8_attention_direct(query, key, value, is_causal, scale, scale_q, scale_k, scaling_method, scale_v)
    if scale_v is not None:
        raise ValueError("a pre-quantised value (scale_v) is only supported on the eager path")
    if not traced and config.attention.force_eager_fallback and scale_q is None:
        # debug comparator: the reference's eager composite - aten quantiser, then the op's aten definition
        dims = [query.dim() - 2, query.dim() - 1] if scaling_method == "head-wise" else query.dim() - 1
        if scale_k is None:
            key, scale_k = _dynamically_quantize_fp8(key, reduction_dim=dims)
        query, scale_q = _dynamically_quantize_fp8(query, reduction_dim=dims)
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

    traced = torch.compiler.is_dynamo_compiling() or not (
        _plain_tensor(query) and _plain_tensor(key) and _plain_tensor(value)) and any(
        is_fake(x) for x in (query, key, value))
    if not traced and not config.attention.force_eager_fallback:
        return ops.attention_native(query, key, value, is_causal=is_causal, scale=scale)  # no dispatcher hop
    return torch.ops.quantum_attn.attention_forward(
        query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal, scale=scale
    )
