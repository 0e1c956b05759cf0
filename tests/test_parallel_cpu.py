"""Multi-GPU host logic on CPU: head sharding, and the sequence ring under gloo with world size 2 and 3 (the kernels
replaced by the oracle through the backend hook; the NCCL path itself is covered by tests/test_ring_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from quantumattention_b200 import parallel


def test_shard_heads_partitions_exactly():
    for H in (1, 3, 8, 24, 32):
        for W in (1, 2, 4, 8):
            ranges = [parallel.shard_heads(H, W, r) for r in range(W)]
            assert ranges[0][0] == 0 and ranges[-1][1] == H
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a1 >= a0
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert parallel.shard_heads(24, 8, 3) == (9, 12)  # BASELINE C2/C4: 24 heads over 8 GPUs -> 3 each
    with pytest.raises(ValueError):
        parallel.shard_heads(8, 2, 2)


def test_ring_block_owner():
    W = 4
    for r in range(W):
        assert sorted(parallel.ring_block_owner(r, s, W) for s in range(W)) == list(range(W))
        assert parallel.ring_block_owner(r, 0, W) == r


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ring_worker(rank, world, port, S_local, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from ring_backends import OracleBackend

        B, H, D = 1, 2, 64
        q, k, v = oracle.make_qkv(B, H, S_local * world, S_local * world, D, seed=3)
        # make the shards' amax differ so that the MAX all-reduce matters
        k = k * torch.linspace(0.5, 2.0, S_local * world).view(1, 1, -1, 1).to(k.dtype)
        sl = slice(rank * S_local, (rank + 1) * S_local)
        res = {}
        loc = [t[:, :, sl].contiguous() for t in (q, k, v)]
        for pv in ("fp8", "16bit"):
            be = OracleBackend()
            out = parallel.ring_fp8_attention(*loc, backend=be, strategy="ring", pv_mode=pv)
            assert [c[0] if isinstance(c, tuple) else c for c in be.calls] == \
                ["attend", "merge_first"] + ["attend", "merge"] * (world - 1)
            # the gather strategy: per head group ONE launch over all keys, no merge; same quantised bytes
            be2 = OracleBackend()
            out_g = parallel.ring_fp8_attention(*loc, backend=be2, strategy="gather", pv_mode=pv, head_groups=2)
            assert be2.calls == [("attend", (B, 1, S_local, D), (B, 1, S_local * world, D))] * H
            assert torch.allclose(out_g.float(), out.float(), atol=2e-2, rtol=2e-2)
            be3 = OracleBackend()
            out_1 = parallel.ring_fp8_attention(*loc, backend=be3, strategy="gather", pv_mode=pv)
            assert len(be3.calls) == 1 and torch.equal(out_1, out_g)  # (tiny problem: one group of all heads)
            res[pv] = (out, out_g)
        assert parallel.default_seq_strategy() in parallel.SEQ_STRATEGIES
        assert parallel.default_seq_transport() in parallel.SEQ_TRANSPORTS
        # head-sharded layout on the same data: rank r computes its heads with a stand-in kernel, gather returns all
        def fake_attn(q_, k_, v_):
            return (q_.float() + k_.float().mean(2, keepdim=True) + v_.float().mean(2, keepdim=True)).to(q_.dtype)
        full = parallel.head_sharded_fp8_attention(q, k, v, gather=(H % world == 0), _attn=fake_attn)
        torch.save({"res": res, "heads": full}, os.path.join(tmp, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_ring_matches_unsharded_oracle(world, tmp_path):
    S_local = 96  # ragged against the kernel's 128-key tiles on purpose
    mp.spawn(_ring_worker, args=(world, _free_port(), S_local, str(tmp_path)), nprocs=world, join=True)
    B, H, D = 1, 2, 64
    S = S_local * world
    q, k, v = oracle.make_qkv(B, H, S, S, D, seed=3)
    k = k * torch.linspace(0.5, 2.0, S).view(1, 1, -1, 1).to(k.dtype)
    # the unsharded path: quantise the whole sequence head-wise, fp64 attention on the dequantised tensors
    q8, sq = oracle.quantize_fp8(q.float().numpy(), "head-wise")
    k8, sk = oracle.quantize_fp8(k.float().numpy(), "head-wise")
    v8, sv = oracle.quantize_fp8(v.float().numpy(), "head-wise")
    refs = {"fp8": oracle.fp8_attention_ref(q8, k8, v8, sq, sk, scale_v=sv),
            "16bit": oracle.fp8_attention_ref(q8, k8, v.float().numpy(), sq, sk)}  # the reference's op: 16-bit V
    for pv, ref in refs.items():
        for i, key in enumerate(("ring", "gather")):
            got = torch.cat([torch.load(os.path.join(tmp_path, f"r{r}.pt"))["res"][pv][i] for r in range(world)], dim=2)
            m = oracle.compare(got.float().numpy(), ref.numpy())
            # bf16 partial results and a bf16 output: two roundings of 2^-9
            assert m["finite"] and m["cos_sim"] > 0.99999 and m["max_abs_over_row_rms"] < 2e-2, (pv, key, m)
    if H % world == 0:
        heads = [torch.load(os.path.join(tmp_path, f"r{r}.pt"))["heads"] for r in range(world)]
        want = (q.float() + k.float().mean(2, keepdim=True) + v.float().mean(2, keepdim=True)).to(q.dtype)
        for h_ in heads:
            assert torch.equal(h_, want)


def test_given_scale_quantiser_matches_plain_head_wise():
    x = oracle.make_qkv(1, 3, 200, 200, 64, seed=5)[0].float().numpy()
    b0, s0 = oracle.quantize_fp8(x, "head-wise")
    s1 = oracle.head_scales(x)
    assert np.array_equal(s0, s1)
    assert np.array_equal(b0, oracle.quantize_with_scale(x, s1))
    # the scale of a whole head is the max of the scales of its sequence shards
    parts = [oracle.head_scales(x[:, :, a:a + 50]) for a in range(0, 200, 50)]
    assert np.array_equal(np.maximum.reduce(parts), s0)


def test_merge_ref_is_exact_split_softmax():
    g = torch.Generator().manual_seed(0)
    s = torch.randn(4, 37, generator=g, dtype=torch.float64) * 3
    v = torch.randn(37, 8, generator=g, dtype=torch.float64)
    full = torch.softmax(s, -1) @ v
    a, b = slice(0, 20), slice(20, 37)
    oa, la = torch.softmax(s[:, a], -1) @ v[a], torch.logsumexp(s[:, a], -1)
    ob, lb = torch.softmax(s[:, b], -1) @ v[b], torch.logsumexp(s[:, b], -1)
    o, l = oracle.merge_ref(oa, la, ob, lb)
    assert torch.allclose(o, full, atol=1e-12) and torch.allclose(l, torch.logsumexp(s, -1), atol=1e-12)
    # an empty side (LSE = -inf) contributes nothing
    o2, l2 = oracle.merge_ref(oa, la, torch.zeros_like(ob), torch.full_like(lb, float("-inf")))
    assert torch.equal(o2, oa) and torch.equal(l2, la)


def test_gated_launch_is_chosen_when_an_sm_gets_enough_ctas(monkeypatch):
    """One gated launch over all heads pays from 8 CTAs per SM per call (C4: N = 2 and 4, not N = 8); QA_SEQ_GATED forces."""
    monkeypatch.delenv("QA_SEQ_GATED", raising=False)
    assert parallel.seq_gated_launch(1, 24, 37800) and parallel.seq_gated_launch(1, 24, 18900)
    assert not parallel.seq_gated_launch(1, 24, 9450)
    monkeypatch.setenv("QA_SEQ_GATED", "1")
    assert parallel.seq_gated_launch(1, 24, 9450)
    monkeypatch.setenv("QA_SEQ_GATED", "0")
    assert not parallel.seq_gated_launch(1, 24, 37800)


def test_head_chunks_fill_whole_waves():
    """Head groups of the gather strategy: contiguous, cover every head once, and sized so a launch fills the SMs in
    whole waves where the shape allows it (C4 over 8 / 4 / 2 ranks: 37 / 74 / 148 CTAs per head on 148 SMs)."""
    for world, want_hc in ((8, 4), (4, 2), (2, 2)):
        ch = parallel.head_chunks(1, 24, 75600 // world)
        assert ch[0] == (0, want_hc) and ch[-1][1] == 24
        assert all(a[1] == b[0] for a, b in zip(ch, ch[1:])) and len(ch) <= 12
        ctas = (ch[0][1] - ch[0][0]) * -(-(75600 // world) // 256)
        assert ctas % 148 == 0
    assert parallel.head_chunks(2, 3, 100) == [(0, 3)]
    assert parallel.head_chunks(1, 1, 10 ** 6) == [(0, 1)]


def test_unknown_sequence_strategy_is_rejected(monkeypatch):
    monkeypatch.setenv("QA_SEQ_STRATEGY", "tree")
    with pytest.raises(ValueError, match="QA_SEQ_STRATEGY"):
        parallel.default_seq_strategy()

