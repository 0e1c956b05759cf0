"""Multi-GPU layouts of the FP8 attention path: one process per GPU, ``torch.distributed`` for the plumbing.

The reference is single-device (its launcher puts a DeviceGuard on q's device and launches on the current stream,
src/quantum_attn/tk/attention.py:417,465); callers such as ParaAttention shard the work around it.  Two layouts:

* **head sharding** - every (batch, head) is an independent problem (the kernel's grid y/z axes,
  src/quantum_attn/tk/attention.py:504; head-wise scales are per (b, h), src/quantum_attn/nn.py:411-412), so ranks
  take contiguous head ranges and there is NO collective: ``shard_heads`` / ``head_sharded_fp8_attention``.
* **sequence ring** for shapes whose single (b, h) problems are too long for one GPU's share of the latency budget
  (BASELINE config 4, S = 75 600): each rank owns S/N tokens of Q, K and V.  Q, K, V are quantised ONCE to e4m3 with
  head scales made global by a single ``all_reduce(MAX)`` (12*B*H bytes) - so the bytes are the ones the unsharded
  call would produce - and the e4m3 K/V blocks (half the bytes of the 16-bit tensors) reach the other ranks either by
  ONE all-gather over NVSwitch that runs under the attention of the local block (default; two launches and one merge
  whatever the world size) or round a neighbour ring with ``batch_isend_irecv`` while the fused kernel attends the
  block already here.  Partial results carry their log-sum-exp and are combined by ``qa_merge_partials``.
  Non-causal only (as the BASELINE config).

The kernels are reached through a small backend object so that the host logic (ring order, buffer rotation, scale
exchange, merge order) can be exercised on CPU with the ``gloo`` backend in tests; the default backend is the
sm_100a library and has no fallback.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _native, ops


# ------------------------------------------------------------------------------------------------ head sharding
def shard_heads(n_heads: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous head range [start, stop) of ``rank``; the first ``n_heads % world_size`` ranks take one extra."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    base, extra = divmod(n_heads, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def head_sharded_fp8_attention(q, k, v, *, is_causal=False, scale=None, scaling_method="head-wise", group=None,
                               gather: bool = False, _attn=None):
    """Run this rank's head range of a replicated [B,H,S,D] problem; no collective unless ``gather`` asks for the
    full output back (one all_gather of the outputs; head counts must then divide evenly)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    Hq, Hkv = q.shape[1], k.shape[1]
    if Hq % Hkv:
        raise ValueError(f"Expect Hq to be a multiple of Hkv but got Hq={Hq} and Hkv={Hkv}.")
    lo, hi = shard_heads(Hkv, world, rank)  # shard KV heads so that GQA groups stay together
    g = Hq // Hkv
    if _attn is None:
        from . import nn as _nn

        def _attn(q_, k_, v_):
            return _nn.fp8_attention(q_, k_, v_, is_causal=is_causal, scale=scale, scaling_method=scaling_method)
    if hi > lo:
        out = _attn(q[:, lo * g:hi * g], k[:, lo:hi], v[:, lo:hi])
    else:
        out = q.new_empty((q.shape[0], 0, q.shape[2], q.shape[3]))
    if not gather:
        return out
    if Hkv % world:
        raise ValueError("gather=True needs the head count to divide evenly over the ranks")
    parts = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(parts, out.contiguous(), group=group)
    return torch.cat(parts, dim=1)


# ------------------------------------------------------------------------------------------------ sequence ring
class NativeBackend:
    """The sm_100a kernels (include/qattn.h).  Raises if the library or the device is missing."""

    def local_scales(self, tensors: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        _, scales = _native.quantize_fp8(list(tensors), _native.QA_SCALE_HEAD_AMAX_ONLY)
        return scales

    def quantize(self, tensors: Sequence[torch.Tensor], scales: Sequence[torch.Tensor]) -> List[torch.Tensor]:
        outs, _ = _native.quantize_fp8(list(tensors), _native.QA_SCALE_HEAD_GIVEN, scales=list(scales))
        return outs

    def attend(self, q8, k8, v8, sq, sk, sv, sm_scale, p_mode, out_dtype):
        return _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=_native.QA_SCALE_HEAD, is_causal=False,
                                    sm_scale=sm_scale, p_mode=p_mode, out_dtype=out_dtype, return_lse=True)

    def merge(self, o_acc, lse_acc, o_new, lse_new, first, out=None):
        _native.merge_partials(o_acc, lse_acc, o_new, lse_new, first=first, out=out)


def _ring_exchange(send_buf: torch.Tensor, recv_buf: torch.Tensor, group) -> list:
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nxt = dist.get_global_rank(group, (rank + 1) % world) if group is not None else (rank + 1) % world
    prv = dist.get_global_rank(group, (rank - 1) % world) if group is not None else (rank - 1) % world
    ops_ = [dist.P2POp(dist.isend, send_buf, nxt, group), dist.P2POp(dist.irecv, recv_buf, prv, group)]
    return dist.batch_isend_irecv(ops_)


SEQ_STRATEGIES = ("ring", "gather")


def seq_pv_mode() -> str:
    """P mode of the sequence-sharded path when the caller names none: the configured one, except that the library
    default "16bit" (V stays 16-bit) cannot apply - e4m3 K/V on the wire is the point of this path - and maps to
    "fp8"; ask for "fp8_hilo" to stay inside the 2e-2 max-abs bound."""
    from . import config

    return "fp8" if config.attention.pv_mode == "16bit" else config.attention.pv_mode


def default_seq_strategy() -> str:
    """``QA_SEQ_STRATEGY`` (ring | gather), else the measured choice."""
    s = os.environ.get("QA_SEQ_STRATEGY", "gather")
    if s not in SEQ_STRATEGIES:
        raise ValueError(f"QA_SEQ_STRATEGY must be one of {SEQ_STRATEGIES} but got {s!r}")
    return s


def _gather_blocks(kv_loc: torch.Tensor, world: int, group):
    """Start the all-gather of every rank's [2,B,H,S,D] e4m3 K/V block; returns ([world,2,B,H,S,D] bytes, work)."""
    # (the gloo backend of the CPU tests wants the output as a dim-0 concatenation of the inputs)
    flat = torch.empty((world * kv_loc.shape[0],) + tuple(kv_loc.shape[1:]), dtype=kv_loc.dtype, device=kv_loc.device)
    work = dist.all_gather_into_tensor(flat, kv_loc, group=group, async_op=True)
    return flat.view((world,) + tuple(kv_loc.shape)), work


def _concat_other_blocks(kv_all: torch.Tensor, rank: int) -> torch.Tensor:
    """[world,2,B,H,S,D] gathered bytes -> [2,B,H,(world-1)*S,D]: the keys / values of every OTHER rank laid end to
    end per head (key order is immaterial to non-causal attention).  Two strided copies of 8-byte words."""
    world, two, B, H, S, D = kv_all.shape
    dst = torch.empty((two, B, H, (world - 1) * S, D), dtype=kv_all.dtype, device=kv_all.device)
    wide = torch.int64 if D % 8 == 0 else kv_all.dtype
    src6 = kv_all.view(wide).permute(1, 2, 3, 0, 4, 5)  # [2,B,H,world,S,D/8]
    dst6 = dst.view(wide).view(two, B, H, world - 1, S, -1)
    if rank > 0:
        dst6[:, :, :, :rank].copy_(src6[:, :, :, :rank])
    if rank < world - 1:
        dst6[:, :, :, rank:].copy_(src6[:, :, :, rank + 1:])
    return dst


def ring_fp8_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: Optional[float] = None,
                       pv_mode: Optional[str] = None, group=None, backend=None,
                       strategy: Optional[str] = None) -> torch.Tensor:
    """Non-causal FP8 attention over a sequence sharded across the ranks of ``group``.

    q, k, v: this rank's [B, H, S_local, D] 16-bit slices (equal S_local on every rank); returns the [B, H, S_local, D]
    output rows of the local queries against the keys/values of ALL ranks.  With world size 1 this is exactly
    ``fp8_attn_func(q, k, v)`` in the chosen P mode (``seq_pv_mode()`` when none is given).

    ``strategy``: how the other ranks' e4m3 K/V reach this one.  "ring": world - 1 neighbour exchanges, one kernel launch
    and one merge per block.  "gather": ONE all-gather over NVSwitch (every GPU has full bandwidth to every peer, so
    nothing is won by forwarding hop by hop) that overlaps the attention of the local block, then ONE launch over all
    the other ranks' keys and one merge - two launches whatever the world size.  Same quantised bytes either way.
    """
    be = backend if backend is not None else NativeBackend()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if q.dim() != 4 or k.shape != v.shape or q.shape[:2] != k.shape[:2] or q.shape[3] != k.shape[3]:
        raise ValueError("ring_fp8_attention: q, k, v must be [B,H,S_local,D] with equal B, H, D (no GQA)")
    if pv_mode is None:
        pv_mode = seq_pv_mode()
    p_mode = ops.pv_mode_code(pv_mode)
    if p_mode == _native.QA_P_16BIT:
        raise ValueError("ring_fp8_attention moves e4m3 K/V blocks: pv_mode must be 'fp8' or 'fp8_hilo'")
    strategy = default_seq_strategy() if strategy is None else strategy
    if strategy not in SEQ_STRATEGIES:
        raise ValueError(f"strategy must be one of {SEQ_STRATEGIES} but got {strategy!r}")
    B, H, S, D = q.shape
    sm_scale = (1.0 / math.sqrt(D)) if scale is None else float(scale)

    # 1. head scales of the WHOLE sequence: local scales, one MAX all-reduce (scale is monotone in amax)
    scales = torch.stack(be.local_scales([q, k, v]))  # [3, B, H] fp32
    if world > 1:
        dist.all_reduce(scales, op=dist.ReduceOp.MAX, group=group)
    sq, sk, sv = scales[0], scales[1], scales[2]
    # 2. quantise once; K and V share one buffer so a ring step is one send and one receive
    q8, k8, v8 = be.quantize([q, k, v], [sq, sk, sv])
    kv = [torch.stack((k8.view(torch.uint8), v8.view(torch.uint8))), None]  # [2, B, H, S, D] bytes
    if world > 1 and strategy == "ring":
        kv[1] = torch.empty_like(kv[0])

    out = torch.empty_like(q)
    o_acc = torch.empty((B, H, S, D), dtype=torch.float32, device=q.device) if world > 1 else None
    lse_acc = torch.empty((B, H, S), dtype=torch.float32, device=q.device)
    if strategy == "gather" and world > 1:
        kv_all, work = _gather_blocks(kv[0], world, group)
        # the local block needs nothing from the wire: attend it while the gather runs
        o_new, lse_new = be.attend(q8, k8, v8, sq, sk, sv, sm_scale, p_mode, q.dtype)
        be.merge(o_acc, lse_acc, o_new, lse_new, True, None)
        work.wait()
        rest = _concat_other_blocks(kv_all, rank)
        del kv_all
        o_new, lse_new = be.attend(q8, rest[0].view(torch.float8_e4m3fn), rest[1].view(torch.float8_e4m3fn), sq, sk, sv,
                                   sm_scale, p_mode, q.dtype)
        be.merge(o_acc, lse_acc, o_new, lse_new, False, out)
        return out
    cur = 0
    for step in range(world):
        last = step == world - 1
        # 3. start moving the block we hold to the next rank, then compute on it (the transfer only reads it)
        reqs = _ring_exchange(kv[cur], kv[cur ^ 1], group) if not last else []
        k_blk = kv[cur][0].view(torch.float8_e4m3fn)
        v_blk = kv[cur][1].view(torch.float8_e4m3fn)
        o_new, lse_new = be.attend(q8, k_blk, v_blk, sq, sk, sv, sm_scale, p_mode, q.dtype)
        # 4. fold the partial result in; the last step writes the 16-bit output directly
        be.merge(o_acc, lse_acc, o_new, lse_new, step == 0, out if last else None)
        for r in reqs:
            r.wait()
        cur ^= 1
    return out


# the function predates the all-gather strategy; this is the name that says what it does
sequence_sharded_fp8_attention = ring_fp8_attention


def ring_block_owner(rank: int, step: int, world: int) -> int:
    """Rank whose K/V block ``rank`` holds at ring step ``step`` (blocks travel to rank + 1 each step)."""
    return (rank - step) % world
