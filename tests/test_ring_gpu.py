"""GPU parity of the pieces the sequence ring adds: the (O, LSE) merge kernel, the split quantiser (local scales ->
MAX -> quantise with given scales), ring_fp8_attention on one rank, and - when the box has two GPUs - the NCCL ring
itself against the unsharded kernel.  Everything goes through the C ABI (include/qattn.h)."""
import math
import os
import socket

import numpy as np
import pytest
import torch

import oracle
import quantum_attn
from quantumattention_b200 import _native, parallel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [64, 128, 256])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_merge_kernel_matches_oracle(D, dtype):
    g = torch.Generator().manual_seed(D)
    rows = 1000  # not a multiple of the rows a CTA covers
    o_acc = torch.randn(rows, D, generator=g)
    lse_acc = torch.randn(rows, generator=g) * 4
    o_new = torch.randn(rows, D, generator=g).to(dtype)
    lse_new = torch.randn(rows, generator=g) * 4
    lse_acc[3] = float("-inf")                         # accumulator side empty
    lse_new[5] = float("-inf")                         # new side empty
    lse_acc[7] = lse_new[7] = float("-inf")            # both empty
    ref_o, ref_l = oracle.merge_ref(o_acc.double(), lse_acc.double(), o_new.double(), lse_new.double())
    a, la = o_acc.cuda(), lse_acc.cuda()
    _native.merge_partials(a, la, o_new.cuda(), lse_new.cuda(), first=False)
    torch.cuda.synchronize()
    assert torch.allclose(a.cpu().double(), ref_o, atol=2e-6, rtol=2e-6)
    assert torch.allclose(la.cpu().double()[torch.isfinite(ref_l)], ref_l[torch.isfinite(ref_l)], atol=1e-5)
    assert torch.equal(torch.isinf(la.cpu()), torch.isinf(ref_l))
    assert torch.equal(a.cpu()[7], torch.zeros(D))
    # last-step form: 16-bit output instead of the fp32 accumulator
    a2, la2, out = o_acc.cuda(), lse_acc.cuda(), torch.empty(rows, D, dtype=dtype, device="cuda")
    _native.merge_partials(a2, la2, o_new.cuda(), lse_new.cuda(), first=False, out=out)
    torch.cuda.synchronize()
    assert torch.equal(a2.cpu(), o_acc)  # untouched
    assert torch.equal(out.cpu(), a.to(dtype).cpu())
    # first-step form: plain copy
    a3, la3 = torch.full((rows, D), 7.0, device="cuda"), torch.full((rows,), 7.0, device="cuda")
    _native.merge_partials(a3, la3, o_new.cuda(), lse_new.cuda(), first=True)
    torch.cuda.synchronize()
    assert torch.equal(a3.cpu(), o_new.float()) and torch.equal(la3.cpu(), lse_new)


def test_split_quantiser_is_byte_exact():
    q, k, v = oracle.make_qkv(2, 3, 700, 700, 128, seed=11)
    xs = [q.cuda(), k.cuda(), v.cuda()]
    (a8, b8, c8), (sa, sb, sc) = _native.quantize_fp8(xs, _native.QA_SCALE_HEAD)
    _, scales = _native.quantize_fp8(xs, _native.QA_SCALE_HEAD_AMAX_ONLY)
    for s_, t_ in zip(scales, (sa, sb, sc)):
        assert torch.equal(s_, t_)
    outs, _ = _native.quantize_fp8(xs, _native.QA_SCALE_HEAD_GIVEN, scales=scales)
    for o_, t_ in zip(outs, (a8, b8, c8)):
        assert torch.equal(o_.view(torch.uint8), t_.view(torch.uint8))
    # shards of the sequence: MAX of the shard scales is the scale of the whole head, and the bytes follow
    parts = [_native.quantize_fp8([x[:, :, a:a + 175].contiguous() for x in xs], _native.QA_SCALE_HEAD_AMAX_ONLY)[1]
             for a in range(0, 700, 175)]
    glob = [torch.stack([p[i] for p in parts]).amax(0) for i in range(3)]
    for s_, t_ in zip(glob, (sa, sb, sc)):
        assert torch.equal(s_, t_)
    sh = [x[:, :, 175:350].contiguous() for x in xs]
    outs, _ = _native.quantize_fp8(sh, _native.QA_SCALE_HEAD_GIVEN, scales=glob)
    assert torch.equal(outs[1].view(torch.uint8), b8.view(torch.uint8)[:, :, 175:350])
    ref = oracle.quantize_with_scale(sh[0].float().cpu().numpy(), glob[0].cpu().numpy())
    assert np.array_equal(outs[0].view(torch.uint8).cpu().numpy(), ref)


@pytest.mark.parametrize("pv", ["fp8", "fp8_hilo"])
def test_key_blocks_plus_merge_equal_one_launch(pv):
    """Attend to 3 ragged key blocks separately, merge with the kernel: same result as one launch over all keys."""
    B, H, S, D = 1, 4, 1000, 128
    q, k, v = oracle.make_qkv(B, H, S, S, D, seed=2)
    (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q.cuda(), k.cuda(), v.cuda()], _native.QA_SCALE_HEAD)
    pm = {"fp8": _native.QA_P_E4M3, "fp8_hilo": _native.QA_P_E4M3_HILO}[pv]
    kw = dict(scale_mode=_native.QA_SCALE_HEAD, is_causal=False, sm_scale=1 / math.sqrt(D), p_mode=pm,
              out_dtype=torch.bfloat16)
    full, lse_full = _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, return_lse=True, **kw)
    o_acc = torch.empty(B, H, S, D, device="cuda")
    lse_acc = torch.empty(B, H, S, device="cuda")
    out = torch.empty(B, H, S, D, dtype=torch.bfloat16, device="cuda")
    cuts = [0, 300, 812, 1000]
    for i, (a, b) in enumerate(zip(cuts, cuts[1:])):
        o, l = _native.fp8_attn_fwd(q8, k8[:, :, a:b].contiguous(), v8[:, :, a:b].contiguous(), sq, sk, sv,
                                    return_lse=True, **kw)
        _native.merge_partials(o_acc, lse_acc, o, l, first=(i == 0), out=out if b == S else None)
    torch.cuda.synchronize()
    # single-e4m3 mode: l is the row sum of the QUANTISED probabilities (accumulated by the tensor core, the same
    # weights that normalise O), and how a probability rounds depends on the block's own running maximum
    lse_tol = 1e-2 if pv == "fp8" else 2e-3
    assert torch.allclose(lse_acc, lse_full, atol=lse_tol), (lse_acc - lse_full).abs().max()
    ref = oracle.fp8_attention_ref(q8.view(torch.uint8).cpu().numpy(), k8.view(torch.uint8).cpu().numpy(),
                                   v8.view(torch.uint8).cpu().numpy(), sq.cpu().numpy(), sk.cpu().numpy(),
                                   scale_v=sv.cpu().numpy())
    m_full = oracle.compare(full.float().cpu().numpy(), ref.numpy())
    m_ring = oracle.compare(out.float().cpu().numpy(), ref.numpy())
    assert m_ring["cos_sim"] >= 0.999 and m_ring["rmse"] < 1e-2, m_ring
    # each partial result is rounded to bf16 once more than the single launch
    assert m_ring["rmse"] < 2.0 * m_full["rmse"] + 1e-5, (m_ring, m_full)
    # LSE itself against the oracle
    _, lse_ref = oracle.attention_block_ref(q8.view(torch.uint8).cpu().numpy(), k8.view(torch.uint8).cpu().numpy(),
                                            v8.view(torch.uint8).cpu().numpy(), sq.cpu().numpy(), sk.cpu().numpy(),
                                            sv.cpu().numpy())
    assert torch.allclose(lse_full.cpu().double(), lse_ref, atol=lse_tol)


@pytest.mark.parametrize("pv", ["16bit", "fp8", "fp8_hilo"])
def test_ring_single_rank_is_fp8_attn_func(pv):
    q, k, v = (t.cuda() for t in oracle.make_qkv(1, 4, 640, 640, 128, seed=4))
    a = parallel.ring_fp8_attention(q, k, v, pv_mode=pv)
    with quantum_attn.config.patch({"attention.pv_mode": pv}):
        b = quantum_attn.fp8_attn_func(q, k, v)
        c = parallel.ring_fp8_attention(q, k, v)  # no mode named: the configured one
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(c, b)
    assert parallel.seq_pv_mode() == quantum_attn.config.attention.pv_mode == "16bit"  # default: within the 2e-2 bound


def test_head_group_launches_over_gathered_layout_equal_one_launch():
    """What the gather strategy does on a rank, minus the wire: K / V blocks laid end to end per head in a
    [B,H,world*S,D] buffer, heads attended group by group through strided views, output written head range by head
    range - bit-identical to one launch over everything."""
    B, H, S, D, world = 1, 6, 300, 128, 3
    q, k, v = (t.cuda() for t in oracle.make_qkv(B, H, S * world, S * world, D, seed=6))
    (q8, k8), (sq, sk) = _native.quantize_fp8([q, k], _native.QA_SCALE_HEAD)
    kw = dict(scale_mode=_native.QA_SCALE_HEAD, is_causal=False, sm_scale=1 / math.sqrt(D), p_mode=_native.QA_P_16BIT,
              out_dtype=torch.bfloat16)
    whole = _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, **kw)
    out = torch.empty_like(whole)
    for lo, hi in [(0, 2), (2, 4), (4, 6)]:
        _native.fp8_attn_fwd(q8[:, lo:hi], k8[:, lo:hi], v[:, lo:hi], sq[:, lo:hi], sk[:, lo:hi], None, out=out[:, lo:hi], **kw)
    assert torch.equal(out, whole)
    # B > 1: head ranges are strided views (batch stride spans all heads) and go through the tensor maps as they are
    q2, k2, v2 = (torch.cat([t, t.flip(2)], 0) for t in (q, k, v))
    (q8, k8), (sq, sk) = _native.quantize_fp8([q2, k2], _native.QA_SCALE_HEAD)
    whole = _native.fp8_attn_fwd(q8, k8, v2, sq, sk, None, **kw)
    part = _native.fp8_attn_fwd(q8[:, 1:3], k8[:, 1:3], v2[:, 1:3], sq[:, 1:3], sk[:, 1:3], None, **kw)
    assert not q8[:, 1:3].is_contiguous() and torch.equal(part, whole[:, 1:3])


def test_gated_launch_reads_a_head_group_only_behind_its_flags():
    """qa_fp8_attn_fwd_gated: the kernel is launched while K / V of the head groups are still being written on ANOTHER
    stream; each group's flags are set behind its copies.  The result must be that of the plain launch on the final
    bytes (an early read would see the zero-filled buffers), and the launch must have lasted as long as the delay."""
    B, H, S, D = 1, 4, 640, 128
    q, k, v = (t.cuda() for t in oracle.make_qkv(B, H, S, S, D, seed=12))
    (q8, k8), (sq, sk) = _native.quantize_fp8([q, k], _native.QA_SCALE_HEAD)
    kw = dict(scale_mode=_native.QA_SCALE_HEAD, is_causal=False, sm_scale=1 / math.sqrt(D), p_mode=_native.QA_P_16BIT,
              out_dtype=torch.bfloat16)
    ref = _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, **kw)
    main, side = torch.cuda.current_stream(), torch.cuda.Stream()
    for rep in range(2):  # (second round: flags re-zeroed on the launch stream, as the sequence-sharded path does)
        k_buf, v_buf = torch.zeros_like(k8.view(torch.uint8)).view(k8.dtype), torch.zeros_like(v)
        flags = torch.zeros(4, dtype=torch.int32, device="cuda")  # 2 gates (2 heads each) x 2 flags
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            for g in (1, 0):  # the LATER head group first: order of arrival is free
                torch.cuda._sleep(20_000_000)  # ~10 ms
                k_buf[:, 2 * g:2 * g + 2].copy_(k8[:, 2 * g:2 * g + 2])
                v_buf[:, 2 * g:2 * g + 2].copy_(v[:, 2 * g:2 * g + 2])
                _native.set_flag(flags, 2 * g, side.cuda_stream)
                _native.set_flag(flags, 2 * g + 1, side.cuda_stream)
        e0.record(main)
        out = _native.fp8_attn_fwd(q8, k_buf, v_buf, sq, sk, None, gate=(flags, 2, 2), **kw)
        e1.record(main)
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
        assert e0.elapsed_time(e1) > 10.0, e0.elapsed_time(e1)  # it waited for the second group's flags
        assert flags.tolist() == [0x01010101] * 4
    with pytest.raises(ValueError):
        _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, gate=(flags[:1], 2, 2), **kw)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        # 1184 = 9 * 128 + 32: ragged against the 128-key tiles, but a multiple of the 32 rows a softmax warp owns, so a
        # rank's rows keep their warp (and with it the warp-wide lazy-rescale decisions) of the unsharded call
        B, H, S_local, D = 1, 4, 1184, 128
        S = S_local * world
        q, k, v = oracle.make_qkv(B, H, S, S, D, seed=9)
        sl = slice(rank * S_local, (rank + 1) * S_local)
        loc = [t[:, :, sl].contiguous().cuda() for t in (q, k, v)]
        res, peer_error = {}, None
        for pv in ("16bit", "fp8", "fp8_hilo"):
            with quantum_attn.config.patch({"attention.pv_mode": pv}):
                whole = quantum_attn.fp8_attn_func(q.cuda(), k.cuda(), v.cuda())[:, :, sl]
            for strategy, transport, gated in (("gather", "nccl", "1"), ("gather", "peer", "1"), ("gather", "peer", "0"),
                                               ("ring", "nccl", "1")):
                if transport == "peer" and peer_error is not None:
                    continue
                os.environ["QA_SEQ_GATED"] = gated  # peer transport: one gated launch over all heads / one per group
                try:
                    for rep in range(3):  # repeated calls: the peer transport alternates its two send slots
                        out = parallel.ring_fp8_attention(*loc, pv_mode=pv, strategy=strategy, transport=transport,
                                                          head_groups=2)
                    torch.cuda.synchronize()
                except Exception as e:  # symmetric memory may be unavailable on a box; NCCL must work
                    if transport != "peer":
                        raise
                    peer_error = repr(e)[:300]
                    continue
                res[(pv, strategy, transport + ("" if gated == "1" else "-ungated"))] = (out.cpu(), whole.cpu())
        torch.save({"res": res, "peer_error": peer_error}, os.path.join(tmp, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_nccl_sequence_sharding_matches_unsharded_kernel(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        d = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        res = d["res"]
        if d["peer_error"]:
            print("peer transport unavailable:", d["peer_error"])
        for (pv, strategy, transport), (a, b) in res.items():
            m = oracle.compare(a.float().numpy(), b.float().numpy())
            if strategy == "gather":
                # one launch over the same quantised bytes in the same key order as the unsharded call: identical
                assert torch.equal(a, b), (pv, strategy, transport, m)
            elif pv == "fp8":
                # P is rounded to e4m3 relative to a running maximum that depends on the key-block partition, so two
                # "fp8" results differ by two independent P roundings
                assert m["cos_sim"] > 0.999 and m["max_abs_over_row_rms"] < 0.4, (pv, strategy, m)
            else:
                # hi+lo / 16-bit P: the difference collapses to the bf16 rounding of the partial results (the largest
                # seen is one bf16 ulp of an output element, 2^-10 here: 0.03-0.04 of the row RMS)
                assert m["cos_sim"] > 0.99999 and m["max_abs_over_row_rms"] < 0.06, (pv, strategy, m)
        assert ("16bit", "gather", "nccl") in res and ("16bit", "ring", "nccl") in res
