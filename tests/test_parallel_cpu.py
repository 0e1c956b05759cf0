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
        be = OracleBackend()
        out = parallel.ring_fp8_attention(q[:, :, sl].contiguous(), k[:, :, sl].contiguous(),
                                          v[:, :, sl].contiguous(), backend=be, strategy="ring")
        assert be.calls == ["attend", "merge_first"] + ["attend", "merge"] * (world - 1)
        # the all-gather layout: two launches whatever the world size, same quantised bytes
        be2 = OracleBackend()
        out_g = parallel.ring_fp8_attention(q[:, :, sl].contiguous(), k[:, :, sl].contiguous(),
                                            v[:, :, sl].contiguous(), backend=be2, strategy="gather")
        assert be2.calls == ["attend", "merge_first", "attend", "merge"]
        assert torch.allclose(out_g.float(), out.float(), atol=2e-2, rtol=2e-2)
        assert parallel.default_seq_strategy() in parallel.SEQ_STRATEGIES
        # head-sharded layout on the same data: rank r computes its heads with a stand-in kernel, gather returns all
        def fake_attn(q_, k_, v_):
            return (q_.float() + k_.float().mean(2, keepdim=True) + v_.float().mean(2, keepdim=True)).to(q_.dtype)
        full = parallel.head_sharded_fp8_attention(q, k, v, gather=(H % world == 0), _attn=fake_attn)
        torch.save({"out": out, "out_gather": out_g, "heads": full}, os.path.join(tmp, f"r{rank}.pt"))
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
    ref = oracle.fp8_attention_ref(q8, k8, v8, sq, sk, scale_v=sv)
    for key in ("out", "out_gather"):
        got = torch.cat([torch.load(os.path.join(tmp_path, f"r{r}.pt"))[key] for r in range(world)], dim=2)
        m = oracle.compare(got.float().numpy(), ref.numpy())
        # bf16 partial results and a bf16 output: two roundings of 2^-9
        assert m["finite"] and m["cos_sim"] > 0.99999 and m["max_abs_over_row_rms"] < 2e-2, (key, m)
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


@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (4, 2), (8, 0), (8, 7)])
def test_concat_other_blocks_layout(world, rank):
    """[world,2,B,H,S,D] gathered bytes -> per head, the blocks of every other rank end to end in rank order."""
    B, H, S, D = 1, 3, 5, 64
    g = torch.Generator().manual_seed(world * 10 + rank)
    kv_all = torch.randint(0, 256, (world, 2, B, H, S, D), dtype=torch.uint8, generator=g)
    got = parallel._concat_other_blocks(kv_all, rank)
    want = torch.cat([kv_all[r] for r in range(world) if r != rank], dim=3)
    assert got.shape == (2, B, H, (world - 1) * S, D) and torch.equal(got, want)


def test_unknown_sequence_strategy_is_rejected(monkeypatch):
    monkeypatch.setenv("QA_SEQ_STRATEGY", "tree")
    with pytest.raises(ValueError, match="QA_SEQ_STRATEGY"):
        parallel.default_seq_strategy()

