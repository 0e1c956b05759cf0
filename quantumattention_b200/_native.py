"""ctypes binding of the C ABI in include/qattn.h.  PyTorch only supplies device memory and the current stream.

There is NO fallback: if the shared library is missing or the device is not sm_100 the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional, Sequence, Tuple

import torch

from . import build as _build

QA_DT_BF16, QA_DT_FP16, QA_DT_E4M3 = 0, 1, 2
QA_SCALE_HEAD, QA_SCALE_TOKEN, QA_SCALE_HEAD_TWO_PASS, QA_SCALE_HEAD_AMAX_ONLY, QA_SCALE_HEAD_GIVEN = 0, 1, 2, 3, 4
QA_WS_PERSISTENT = 0x100  # qa_quantize_fp8: the workspace was zeroed once and is reused (include/qattn.h)
QA_P_E4M3, QA_P_E4M3_HILO, QA_P_16BIT = 0, 1, 2
ABI_VERSION = 4

EXPORTED_SYMBOLS = (
    "qa_abi_version",
    "qa_last_error",
    "qa_device_supported",
    "qa_quantize_workspace_floats",
    "qa_quantize_fp8",
    "qa_fp8_attn_fwd",
    "qa_attn_fwd",
    "qa_merge_partials",
    "qa_last_launch_count",
)

_lib = None
_lock = threading.Lock()

# bookkeeping for bench.py: kernels launched through this binding, and an optional CUDA-event timer around the
# attention kernel (events are recorded on the launching stream; a few hundred ns each)
launch_total = 0
attn_events = None  # set to a list to collect (start_event, stop_event) pairs
quant_events = None  # the same for the quantiser launches


class NativeError(RuntimeError):
    pass


def lib_path() -> str:
    # QA_NATIVE_LIB: developer override to A/B kernel variants built by scripts/build_variant.sh
    return os.environ.get("QA_NATIVE_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent and nvcc is available).  Raises if it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise NativeError(f"{path} is missing; run `python -m quantumattention_b200.build`")
            _build.build_library()
        lib = ctypes.CDLL(path)
        lib.qa_abi_version.restype = ctypes.c_int
        lib.qa_last_error.restype = ctypes.c_char_p
        lib.qa_last_launch_count.restype = ctypes.c_int
        lib.qa_device_supported.argtypes = [ctypes.c_int]
        lib.qa_device_supported.restype = ctypes.c_int
        vp = ctypes.c_void_p
        lib.qa_quantize_fp8.argtypes = [
            ctypes.c_int, ctypes.POINTER(vp), ctypes.c_int, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(vp),
            ctypes.POINTER(vp), vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int,
            ctypes.c_int, vp,
        ]
        lib.qa_quantize_fp8.restype = ctypes.c_int
        lib.qa_quantize_workspace_floats.argtypes = [ctypes.c_int] * 4
        lib.qa_quantize_workspace_floats.restype = ctypes.c_size_t
        lib.qa_fp8_attn_fwd.argtypes = [
            vp, vp, vp, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, ctypes.c_int, vp,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
            ctypes.c_float, ctypes.c_int, vp,
        ]
        lib.qa_fp8_attn_fwd.restype = ctypes.c_int
        lib.qa_attn_fwd.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp] + [ctypes.c_int] * 7 + [ctypes.c_float, vp]
        lib.qa_attn_fwd.restype = ctypes.c_int
        lib.qa_merge_partials.argtypes = [vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_longlong, ctypes.c_int,
                                          ctypes.c_int, vp]
        lib.qa_merge_partials.restype = ctypes.c_int
        if lib.qa_abi_version() != ABI_VERSION:
            raise NativeError(f"ABI mismatch: library {lib.qa_abi_version()} != binding {ABI_VERSION}")
        _lib = lib
    return _lib


def _check(rc: int, what: str):
    if rc != 0:
        msg = load().qa_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def _dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.bfloat16:
        return QA_DT_BF16
    if dtype == torch.float16:
        return QA_DT_FP16
    if dtype == torch.float8_e4m3fn:
        return QA_DT_E4M3
    raise ValueError(f"unsupported dtype {dtype}")


def last_launch_count() -> int:
    return int(load().qa_last_launch_count())


_ws_cache = {}  # (device index, stream handle) -> fp32 workspace, zero-filled when allocated, reused by every call


def _persistent_workspace(dev: torch.device, n_floats: int) -> torch.Tensor:
    """The quantiser's workspace under the QA_WS_PERSISTENT contract of include/qattn.h: zeroed once, then only ever
    written by the library, one per (device, stream) so that the calls sharing it are stream-ordered.  Saves the
    per-call clear (a memset node and a dependency edge in front of every head-wise quantisation)."""
    with torch.cuda.device(dev):
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        ws = _ws_cache.get(key)
        if ws is None or ws.numel() < n_floats:
            ws = torch.zeros((max(n_floats, 1 << 16),), dtype=torch.float32, device=dev)
            _ws_cache[key] = ws
    return ws


def quantize_fp8(tensors: Sequence[torch.Tensor], scale_mode: int,
                 scales: Optional[Sequence[torch.Tensor]] = None,
                 workspace: Optional[torch.Tensor] = None) -> Tuple[list, list]:
    """Quantise 1-3 CUDA tensors [B,H,S_i,D] (bf16/fp16, same B,H,D,dtype) in one launch pair.

    Returns ([e4m3 tensors], [fp32 scales: [B,H] head-wise or [B,H,S_i] token-wise]).
    ``workspace``: optional caller-owned scratch (plain contract: contents ignored); by default a per-(device, stream)
    workspace under the QA_WS_PERSISTENT contract is used, which saves the per-call clear.
    ``QA_SCALE_HEAD_AMAX_ONLY`` returns ([], scales) without quantising; ``QA_SCALE_HEAD_GIVEN`` quantises with the
    fp32 [B,H] ``scales`` passed in.
    """
    lib = load()
    n = len(tensors)
    t0 = tensors[0]
    B, H, _, D = t0.shape
    dev = t0.device
    xs = []
    for t in tensors:
        if t.device != dev or t.dtype != t0.dtype or t.dim() != 4 or t.shape[0] != B or t.shape[1] != H or t.shape[3] != D:
            raise ValueError("quantize_fp8: tensors must share device, dtype, B, H and D")
        if t.stride(3) != 1 or any(s % 8 for s in t.stride()[:3]) or t.data_ptr() % 16:
            t = t.contiguous()
        xs.append(t)
    amax_only = scale_mode == QA_SCALE_HEAD_AMAX_ONLY
    outs = [] if amax_only else [torch.empty(t.shape, dtype=torch.float8_e4m3fn, device=dev) for t in xs]
    if scale_mode == QA_SCALE_HEAD_GIVEN:
        if scales is None or len(scales) != n:
            raise ValueError("quantize_fp8: QA_SCALE_HEAD_GIVEN needs one [B,H] fp32 scale tensor per input")
        scales = [s_.to(device=dev, dtype=torch.float32).reshape(B, H).contiguous() for s_ in scales]
        ws_ptr = None
    elif scale_mode in (QA_SCALE_HEAD, QA_SCALE_HEAD_TWO_PASS, QA_SCALE_HEAD_AMAX_ONLY):
        scales = [torch.empty((B, H), dtype=torch.float32, device=dev) for _ in xs]
        n_ws = int(lib.qa_quantize_workspace_floats(B, H, max(t.shape[2] for t in xs), D))
        if workspace is not None:  # caller's scratch, contents ignored (cleared by the call)
            if workspace.dtype != torch.float32 or workspace.numel() < n_ws or workspace.device != dev:
                raise ValueError(f"quantize_fp8: workspace must be >= {n_ws} fp32 elements on {dev}")
            ws = workspace
        elif os.environ.get("QA_NO_PERSISTENT_WS"):  # developer A/B switch: per-call scratch, cleared by the library
            ws = torch.empty((n_ws,), dtype=torch.float32, device=dev)
        else:
            ws = _persistent_workspace(dev, n_ws)
            scale_mode |= QA_WS_PERSISTENT
        ws_ptr = ws.data_ptr()
    else:
        scales = [torch.empty((B, H, t.shape[2]), dtype=torch.float32, device=dev) for t in xs]
        ws_ptr = None
    vp = ctypes.c_void_p
    x_arr = (vp * n)(*[t.data_ptr() for t in xs])
    o_arr = (vp * n)(*[t.data_ptr() for t in outs]) if outs else (vp * n)()
    s_arr = (vp * n)(*[t.data_ptr() for t in scales])
    strides = (ctypes.c_int64 * (4 * n))(*[s for t in xs for s in t.stride()])
    S = (ctypes.c_int * n)(*[t.shape[2] for t in xs])
    with torch.cuda.device(dev):
        tstream = torch.cuda.current_stream(dev)
        stream = tstream.cuda_stream
        if quant_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(tstream)
        rc = lib.qa_quantize_fp8(n, x_arr, _dt_code(t0.dtype), strides, o_arr, s_arr, ws_ptr, B, H, S, D,
                                 scale_mode, stream)
        if quant_events is not None:
            ev1.record(tstream)
            quant_events.append((ev0, ev1))
    _check(rc, "qa_quantize_fp8")
    global launch_total
    launch_total += int(lib.qa_last_launch_count())
    return outs, scales


def fp8_attn_fwd(q8: torch.Tensor, k8: torch.Tensor, v: torch.Tensor, scale_q: torch.Tensor, scale_k: torch.Tensor,
                 scale_v: Optional[torch.Tensor], *, scale_mode: int, is_causal: bool, sm_scale: float, p_mode: int,
                 out_dtype: torch.dtype, return_lse: bool = False):
    """Launch the fused forward kernel.  q8/k8 dense e4m3 [B,H,S,D]; v dense e4m3 (+scale_v) or bf16/fp16."""
    lib = load()
    B, Hq, Sq, D = q8.shape
    Hkv, Skv = k8.shape[1], k8.shape[2]
    dev = q8.device
    q8, k8, v = q8.contiguous(), k8.contiguous(), v.contiguous()
    scale_q = scale_q.to(torch.float32).contiguous()
    scale_k = scale_k.to(torch.float32).contiguous()
    if scale_v is not None:
        scale_v = scale_v.to(torch.float32).contiguous()
    out = torch.empty((B, Hq, Sq, D), dtype=out_dtype, device=dev)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=dev) if return_lse else None
    with torch.cuda.device(dev):
        tstream = torch.cuda.current_stream(dev)
        stream = tstream.cuda_stream
        if attn_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(tstream)
        rc = lib.qa_fp8_attn_fwd(
            q8.data_ptr(), k8.data_ptr(), v.data_ptr(), _dt_code(v.dtype), scale_q.data_ptr(), scale_k.data_ptr(),
            scale_v.data_ptr() if scale_v is not None else None, scale_mode, out.data_ptr(), _dt_code(out_dtype),
            lse.data_ptr() if lse is not None else None, B, Hq, Hkv, Sq, Skv, D, int(bool(is_causal)),
            float(sm_scale), p_mode, stream,
        )
        if attn_events is not None:
            ev1.record(tstream)
            attn_events.append((ev0, ev1))
    _check(rc, "qa_fp8_attn_fwd")
    global launch_total
    launch_total += int(lib.qa_last_launch_count())
    return (out, lse) if return_lse else out


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, is_causal: bool, sm_scale: float,
             return_lse: bool = False):
    """Launch the fused forward kernel on 16-bit q, k, v (all bf16 or all fp16), dense [B,H,S,D]."""
    lib = load()
    B, Hq, Sq, D = q.shape
    Hkv, Skv = k.shape[1], k.shape[2]
    dev = q.device
    q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
    out = torch.empty((B, Hq, Sq, D), dtype=q.dtype, device=dev)
    lse = torch.empty((B, Hq, Sq), dtype=torch.float32, device=dev) if return_lse else None
    with torch.cuda.device(dev):
        tstream = torch.cuda.current_stream(dev)
        if attn_events is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(tstream)
        rc = lib.qa_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), _dt_code(q.dtype), out.data_ptr(),
                             lse.data_ptr() if lse is not None else None, B, Hq, Hkv, Sq, Skv, D,
                             int(bool(is_causal)), float(sm_scale), tstream.cuda_stream)
        if attn_events is not None:
            ev1.record(tstream)
            attn_events.append((ev0, ev1))
    _check(rc, "qa_attn_fwd")
    global launch_total
    launch_total += int(lib.qa_last_launch_count())
    return (out, lse) if return_lse else out


def merge_partials(o_acc: Optional[torch.Tensor], lse_acc: torch.Tensor, o_new: torch.Tensor, lse_new: torch.Tensor,
                   *, first: bool, out: Optional[torch.Tensor] = None) -> None:
    """(o_acc, lse_acc) <- combine with the partial result (o_new, lse_new) of one key block (include/qattn.h).

    o_acc fp32 [..., D] / lse_acc fp32 [...]; o_new 16-bit, lse_new fp32.  With ``out`` (16-bit) the merged rows go
    there instead of into o_acc (the last ring step)."""
    lib = load()
    D = o_new.shape[-1]
    rows = o_new.numel() // D
    for t in (o_acc, lse_acc, o_new, lse_new, out):
        if t is not None and not t.is_contiguous():
            raise ValueError("merge_partials: tensors must be contiguous")
    if lse_acc.dtype != torch.float32 or lse_new.dtype != torch.float32 or (o_acc is not None and o_acc.dtype != torch.float32):
        raise ValueError("merge_partials: accumulators and LSE must be fp32")
    dev = o_new.device
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.qa_merge_partials(o_acc.data_ptr() if o_acc is not None else None, lse_acc.data_ptr(),
                                   o_new.data_ptr(), _dt_code(o_new.dtype), lse_new.data_ptr(),
                                   out.data_ptr() if out is not None else None, rows, D, int(bool(first)), stream)
    _check(rc, "qa_merge_partials")
    global launch_total
    launch_total += int(lib.qa_last_launch_count())
