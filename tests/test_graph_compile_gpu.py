"""The path as callers really drive it: through the registered torch ops, under ``torch.compile``, replayed from CUDA
graphs, with strided ([B,S,H,D]-held) tensors, with K / V quantised once and reused, and from one process that drives
two devices.

Reference behaviour being matched: the op + fake registration (src/quantum_attn/ops.py:98-147), the compiled wrapper
(src/quantum_attn/nn.py:518-539), dense-ification of strided inputs (src/quantum_attn/tk/attention.py:419-421; here
the strides go into the tensor maps instead).  Everything is compared with the eager sm_100a path (bit-exact: same
kernels, same bytes) and the eager path with the oracle.
"""
import ctypes
import math

import numpy as np
import pytest
import torch

import oracle
import quantum_attn
from quantumattention_b200 import _native, ops

pytestmark = pytest.mark.gpu

PV = {"fp8": _native.QA_P_E4M3, "fp8_hilo": _native.QA_P_E4M3_HILO, "16bit": _native.QA_P_16BIT}


def _qkv(B, H, S, D, seed=0, dtype=torch.bfloat16, Hkv=None):
    q, k, v = oracle.make_qkv(B, H, S, S, D, seed=seed, dtype=dtype)
    if Hkv is not None:
        k, v = k[:, :Hkv].contiguous(), v[:, :Hkv].contiguous()
    return q.cuda(), k.cuda(), v.cuda()


# ------------------------------------------------------------------------------------------------ registered ops
@pytest.mark.parametrize("pv", ["fp8", "fp8_hilo", "16bit"])
@pytest.mark.parametrize("causal", [False, True])
def test_fp8_op_native_branch_matches_oracle(pv, causal):
    """torch.ops.quantum_attn.fp8_attention_forward (reference schema) executes the sm_100a kernel, not the aten
    definition, and agrees with the oracle on the same quantised inputs."""
    q, k, v = _qkv(1, 3, 700, 128, seed=21)
    q8, k8, sq, sk = torch.ops.quantum_attn.quantize_qk_fp8(q, k, False)
    before = _native.launch_total
    with quantum_attn.config.patch({"attention.pv_mode": pv}):
        out = torch.ops.quantum_attn.fp8_attention_forward(q8, k8, v, sq, sk, is_causal=causal)
    assert _native.launch_total > before  # the native library launched kernels for this call
    if pv == "16bit":
        vn, svn = v.float().cpu().numpy(), None
        direct = _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, scale_mode=0, is_causal=causal,
                                      sm_scale=1 / math.sqrt(128), p_mode=PV[pv], out_dtype=v.dtype)
        assert torch.equal(out, direct)
    else:
        (v8,), (sv,) = _native.quantize_fp8([v], _native.QA_SCALE_HEAD)
        vn, svn = v8.view(torch.uint8).cpu().numpy(), sv.cpu().numpy()
    ref = oracle.fp8_attention_ref(q8.view(torch.uint8).cpu().numpy(), k8.view(torch.uint8).cpu().numpy(), vn,
                                   sq.cpu().numpy(), sk.cpu().numpy(), scale_v=svn, is_causal=causal)
    m = oracle.compare(out.float().cpu().numpy(), ref.numpy())
    assert m["cos_sim"] >= 0.999, m
    if pv != "fp8":
        assert m["max_abs_over_row_rms"] <= 0.02, m


def test_quantise_ops_are_the_native_quantiser():
    q, k, _ = _qkv(2, 3, 333, 64, seed=5)
    q8, k8, sq, sk = torch.ops.quantum_attn.quantize_qk_fp8(q, k, False)
    rq, rs = oracle.quantize_fp8(q.float().cpu().numpy(), "head-wise")
    assert np.array_equal(q8.view(torch.uint8).cpu().numpy(), rq) and np.array_equal(sq.cpu().numpy(), rs)
    t8, ts = torch.ops.quantum_attn.dynamically_quantize_fp8(k, [-1])
    rk, rks = oracle.quantize_fp8(k.float().cpu().numpy(), "token-wise")
    assert np.array_equal(t8.view(torch.uint8).cpu().numpy(), rk) and np.array_equal(ts.cpu().numpy(), rks)
    q8t, k8t, sqt, skt = torch.ops.quantum_attn.quantize_qk_fp8(q, k, True)
    assert torch.equal(k8t, t8) and torch.equal(skt, ts) and sqt.shape == q.shape[:-1]


def test_opcheck_registered_ops():
    """torch.library.opcheck on every custom op: fake-tensor agreement (shapes, dtypes, strides, devices of the real
    and the fake implementation) for all of them, and the schema test (no hidden mutation / aliasing) where torch can
    run it - its input comparison has no float8 kernels, so for the ops that touch e4m3 tensors the mutation check is
    done here on the bytes."""
    q, k, v = _qkv(1, 2, 256, 128, seed=2)
    q8, k8, sq, sk = torch.ops.quantum_attn.quantize_qk_fp8(q, k, False)
    fake = ("test_faketensor",)
    torch.library.opcheck(torch.ops.quantum_attn.attention_forward.default, (q, k, v), {"is_causal": False},
                          test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(torch.ops.quantum_attn.fp8_attention_forward.default, (q8, k8, v, sq, sk),
                          {"is_causal": True}, test_utils=fake)
    torch.library.opcheck(torch.ops.quantum_attn.dynamically_quantize_fp8.default, (q, [2, 3]), test_utils=fake)
    torch.library.opcheck(torch.ops.quantum_attn.dynamically_quantize_fp8.default, (q, [-1]), test_utils=fake)
    torch.library.opcheck(torch.ops.quantum_attn.quantize_qk_fp8.default, (q, k, False), test_utils=fake)
    torch.library.opcheck(torch.ops.quantum_attn.quantize_qk_fp8.default, (q, k, True), test_utils=fake)
    # mutates_args=(): inputs come back bit-identical, outputs alias none of them
    snap = [t.clone() for t in (q, k, v, q8.view(torch.uint8), k8.view(torch.uint8), sq, sk)]
    out = torch.ops.quantum_attn.fp8_attention_forward(q8, k8, v, sq, sk, is_causal=True)
    a8, b8, sa, sb = torch.ops.quantum_attn.quantize_qk_fp8(q, k, True)
    t8, ts = torch.ops.quantum_attn.dynamically_quantize_fp8(v, [2, 3])
    torch.cuda.synchronize()
    for before, after in zip(snap, (q, k, v, q8.view(torch.uint8), k8.view(torch.uint8), sq, sk)):
        assert torch.equal(before, after)
    ins = {t.data_ptr() for t in (q, k, v, q8, k8, sq, sk)}
    assert not ({t.data_ptr() for t in (out, a8, b8, sa, sb, t8, ts)} & ins)


# ------------------------------------------------------------------------------------------------ torch.compile
def _graph_ops(fn, *args, **kw):
    """Names of the call_function targets dynamo records for fn(*args) (fullgraph), plus the compiled result."""
    seen = []

    def backend(gm, example_inputs):
        seen.extend(str(n.target) for n in gm.graph.nodes if n.op == "call_function")
        return gm.forward

    torch._dynamo.reset()
    out = torch.compile(fn, backend=backend, fullgraph=True)(*args, **kw)
    return seen, out


@pytest.mark.parametrize("method", ["head-wise", "token-wise"])
def test_compiled_graph_calls_the_hand_written_kernels(method):
    """A traced fp8_attn_func is ONE graph (fullgraph) of exactly two custom ops - the native quantiser and the
    reference-named attention op - with no aten arithmetic for Inductor to turn into generated kernels; it launches
    exactly two kernels in the default mode and returns the eager path's bytes."""
    q, k, v = _qkv(1, 4, 640, 128, seed=9)
    fn = quantum_attn.fp8_attn_func if method == "head-wise" else quantum_attn.fp8_token_wise_attn_func
    eager = fn(q, k, v, is_causal=True)
    names, out = _graph_ops(fn, q, k, v, is_causal=True)
    assert sum(n.startswith("quantum_attn.quantize_qk_fp8") for n in names) == 1, names
    assert sum(n.startswith("quantum_attn.fp8_attention_forward") for n in names) == 1, names
    # nothing else but the unpacking of the quantiser's outputs: no aten arithmetic, no device query
    assert all(n.startswith("quantum_attn.") or "getitem" in n for n in names), names
    assert torch.equal(out, eager)
    compiled = torch.compile(fn, backend="aot_eager", fullgraph=True)
    compiled(q, k, v, is_causal=True)
    before = _native.launch_total
    out2 = compiled(q, k, v, is_causal=True)
    assert _native.launch_total - before == 2  # quant (Q and K in one launch) + attn_fwd_kernel
    assert torch.equal(out2, eager)


def test_torch_compile_inductor_fullgraph():
    """The reference's mode of use: torch.compile(..., fullgraph=True) with the default (Inductor) backend around a
    function that calls the public entry points (src/quantum_attn/nn.py:518-539)."""
    q, k, v = _qkv(2, 4, 512, 64, seed=13)

    def block(q, k, v):
        a = quantum_attn.fp8_attn_func(q, k, v, is_causal=False)
        b = quantum_attn.attn_func(q, k, v, is_causal=True)
        t8, ts = quantum_attn.dynamically_quantize_fp8(v, reduction_dim=[2, 3])
        return a, b, t8, ts

    want = block(q, k, v)
    torch._dynamo.reset()
    got = torch.compile(block, fullgraph=True)(q, k, v)
    for w, g in zip(want, got):
        assert torch.equal(w.view(torch.uint8) if w.dtype == torch.float8_e4m3fn else w,
                           g.view(torch.uint8) if g.dtype == torch.float8_e4m3fn else g)


# ------------------------------------------------------------------------------------------------ CUDA graphs
def _capture(fn, *static_inputs):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn(*static_inputs)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn(*static_inputs)
    return g, out


def test_cuda_graph_replay_of_public_call_with_new_data_every_replay():
    """fp8_attn_func captured once, replayed with DIFFERENT data (and different amax, up and down) every time:
    each replay must quantise with its own scales - the rendezvous tag of the single-pass quantiser advances on the
    device - and reproduce the eager result bit for bit."""
    B, H, S, D = 1, 8, 2048, 128
    sq_, sk_, sv_ = (torch.empty((B, H, S, D), dtype=torch.bfloat16, device="cuda") for _ in range(3))
    g0 = torch.Generator(device="cuda").manual_seed(1)
    for t in (sq_, sk_, sv_):
        t.copy_(torch.randn(t.shape, device="cuda", generator=g0).to(torch.bfloat16))
    graph, out = _capture(lambda a, b, c: quantum_attn.fp8_attn_func(a, b, c), sq_, sk_, sv_)
    for i, gain in enumerate([1.0, 0.01, 37.0, 0.3, 0.3, 5.0]):
        gen = torch.Generator(device="cuda").manual_seed(100 + i)
        q, k, v = ((torch.randn((B, H, S, D), device="cuda", generator=gen) * gain).to(torch.bfloat16) for _ in range(3))
        sq_.copy_(q), sk_.copy_(k), sv_.copy_(v)
        graph.replay()
        torch.cuda.synchronize()
        want = quantum_attn.fp8_attn_func(q, k, v)
        assert torch.equal(out, want), (i, gain)


@pytest.mark.parametrize("persistent", [True, False])
def test_cuda_graph_replay_of_c_abi_quantiser(persistent):
    """qa_quantize_fp8 captured straight through the C ABI with a caller-held workspace - under the QA_WS_PERSISTENT
    contract (no per-call clear; the case that used to reuse one host-side tag for every replay) and as plain scratch.
    Replays see data whose amax goes down and up; every replay's bytes must equal the oracle's."""
    lib = _native.load()
    B, H, S, D = 1, 6, 4096, 128
    x = torch.empty((B, H, S, D), dtype=torch.bfloat16, device="cuda")
    x8 = torch.empty((B, H, S, D), dtype=torch.uint8, device="cuda")
    sc = torch.empty((B, H), dtype=torch.float32, device="cuda")
    ws = torch.zeros((int(lib.qa_quantize_workspace_floats(B, H, S, D)),), dtype=torch.float32, device="cuda")
    vp = ctypes.c_void_p
    xa, oa, sa = (vp * 1)(x.data_ptr()), (vp * 1)(x8.data_ptr()), (vp * 1)(sc.data_ptr())
    strides = (ctypes.c_int64 * 4)(*x.stride())
    Sarr = (ctypes.c_int * 1)(S)
    mode = _native.QA_SCALE_HEAD | (_native.QA_WS_PERSISTENT if persistent else 0)

    def call():
        rc = lib.qa_quantize_fp8(1, xa, _native.QA_DT_BF16, strides, oa, sa, ws.data_ptr(), B, H, Sarr, D, mode,
                                 torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.qa_last_error()

    x.copy_(torch.randn(x.shape, device="cuda").to(torch.bfloat16))
    call()  # eager warm-up on the same workspace (advances the generation outside the graph, too)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        call()
    for i, gain in enumerate([1.0, 0.02, 50.0, 50.0, 0.5, 3.0, 0.001]):
        gen = torch.Generator(device="cuda").manual_seed(7 + i)
        x.copy_((torch.randn(x.shape, device="cuda", generator=gen) * gain).to(torch.bfloat16))
        g.replay()
        if i == 3:
            call()  # an eager call between replays shares the workspace and must not disturb them either
        torch.cuda.synchronize()
        r8, rs = oracle.quantize_fp8(x.float().cpu().numpy(), "head-wise")
        assert np.array_equal(sc.cpu().numpy(), rs), (i, gain)
        assert np.array_equal(x8.cpu().numpy(), r8), (i, gain)


# ------------------------------------------------------------------------------------------------ strides
@pytest.mark.parametrize("D", [64, 128, 256])
def test_bshd_held_inputs_are_read_in_place_and_match_dense(D):
    """q, k, v held as [B,S,H,D] (what a fused QKV projection leaves behind) and passed as permuted views: results are
    bit-identical to the dense copies, for the FP8 entry point in every P mode, for pre-quantised q8 / k8 views and
    for the 16-bit entry point."""
    B, H, S = 2, 3, 520
    g = torch.Generator(device="cuda").manual_seed(3)
    raw = [torch.randn((B, S, H, D), device="cuda", generator=g).to(torch.bfloat16) for _ in range(3)]
    qv, kv, vv = (t.permute(0, 2, 1, 3) for t in raw)
    assert not qv.is_contiguous() and _native._in_place_ok(qv)
    qc, kc, vc = (t.contiguous() for t in (qv, kv, vv))
    for pv in ("16bit", "fp8", "fp8_hilo"):
        with quantum_attn.config.patch({"attention.pv_mode": pv}):
            assert torch.equal(quantum_attn.fp8_attn_func(qv, kv, vv, is_causal=True),
                               quantum_attn.fp8_attn_func(qc, kc, vc, is_causal=True)), pv
    assert torch.equal(quantum_attn.attn_func(qv, kv, vv, is_causal=True), quantum_attn.attn_func(qc, kc, vc, is_causal=True))
    assert torch.equal(quantum_attn.fp8_token_wise_attn_func(qv, kv, vv), quantum_attn.fp8_token_wise_attn_func(qc, kc, vc))
    # pre-quantised tensors held [B,S,H,D]
    q8, sq = quantum_attn.dynamically_quantize_fp8(qc, reduction_dim=[2, 3])
    k8, sk = quantum_attn.dynamically_quantize_fp8(kc, reduction_dim=[2, 3])
    q8v = q8.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
    k8v = k8.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)
    assert torch.equal(quantum_attn.fp8_attn_func(q8v, k8v, vv, scale_q=sq, scale_k=sk),
                       quantum_attn.fp8_attn_func(q8, k8, vc, scale_q=sq, scale_k=sk))
    # a view the kernels cannot read in place (rows not 16-byte aligned) is made dense by the glue
    odd = torch.randn((B, H, S, D + 8), device="cuda", generator=g).to(torch.bfloat16)[..., 4:D + 4]
    assert not _native._in_place_ok(odd)
    assert torch.equal(quantum_attn.attn_func(odd, kv, vv), quantum_attn.attn_func(odd.contiguous(), kc, vc))


# ------------------------------------------------------------------------------------------------ K / V reuse
@pytest.mark.parametrize("pv", ["16bit", "fp8", "fp8_hilo"])
@pytest.mark.parametrize("method", ["head-wise", "token-wise"])
def test_kv_quantised_once_and_reused(pv, method):
    q, k, v = _qkv(1, 4, 384, 128, seed=31)
    q2 = (q.float() * 0.7 + 0.1).to(torch.bfloat16)
    fn = quantum_attn.fp8_attn_func
    with quantum_attn.config.patch({"attention.pv_mode": pv}):
        kv = quantum_attn.quantize_kv(k, v, scaling_method=method)
        assert kv.key.dtype == torch.float8_e4m3fn and (kv.scale_v is None) == (pv == "16bit")
        for qq in (q, q2):  # the same quantised K / V serve several query sets
            a = fn(qq, kv.key, kv.value, scale_k=kv.scale_k, scale_v=kv.scale_v, scaling_method=method, is_causal=True)
            b = fn(qq, k, v, scaling_method=method, is_causal=True)
            assert torch.equal(a, b)


def test_scale_shapes_are_validated():
    """Scales whose shape contradicts scaling_method, or the tensor they scale, are rejected instead of being read
    out of bounds / silently misapplied (eager and traced paths agree on the rule: the scales' shape decides)."""
    q, k, v = _qkv(1, 2, 256, 128, seed=3)
    q8h, sqh = quantum_attn.dynamically_quantize_fp8(q, reduction_dim=[2, 3])
    k8h, skh = quantum_attn.dynamically_quantize_fp8(k, reduction_dim=[2, 3])
    q8t, sqt = quantum_attn.dynamically_quantize_fp8(q, reduction_dim=-1)
    k8t, skt = quantum_attn.dynamically_quantize_fp8(k, reduction_dim=-1)
    ok_h = quantum_attn.fp8_attn_func(q8h, k8h, v, scale_q=sqh, scale_k=skh)
    ok_t = quantum_attn.fp8_token_wise_attn_func(q8t, k8t, v, scale_q=sqt, scale_k=skt)
    assert torch.equal(ok_h, quantum_attn.fp8_attn_func(q, k, v))
    assert torch.equal(ok_t, quantum_attn.fp8_token_wise_attn_func(q, k, v))
    with pytest.raises(ValueError, match="contradicts"):
        quantum_attn.fp8_attn_func(q8t, k8t, v, scale_q=sqt, scale_k=skt)  # [B,H,S] scales, head-wise method
    with pytest.raises(ValueError, match="contradicts"):
        quantum_attn.fp8_token_wise_attn_func(q8h, k8h, v, scale_q=sqh, scale_k=skh)  # [B,H] scales, token-wise method
    with pytest.raises(ValueError, match="granularity"):
        quantum_attn.fp8_attn_func(q8h, k8t, v, scale_q=sqh, scale_k=skt)
    with pytest.raises(ValueError, match="matches neither"):
        quantum_attn.fp8_attn_func(q8h, k8h, v, scale_q=sqh[:, :1], scale_k=skh)
    with pytest.raises(ValueError):
        ops.fp8_attention_native(q8h, k8h, v, sqh.flatten()[:1], skh)
    with pytest.raises(ValueError):
        _native.fp8_attn_fwd(q8h, k8h, v, sqh, skh, None, scale_mode=_native.QA_SCALE_TOKEN, is_causal=False,
                             sm_scale=0.1, p_mode=_native.QA_P_16BIT, out_dtype=torch.bfloat16)


# ------------------------------------------------------------------------------------------------ two devices
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_one_process_drives_two_devices():
    """Function attributes and SM counts are per device: the first launch on a second device of the same process
    must opt in to its shared memory there too."""
    outs = []
    for dev in (0, 1, 0, 1):
        q, k, v = (t.to(f"cuda:{dev}") for t in oracle.make_qkv(1, 2, 512, 512, 128, seed=4))
        outs.append(quantum_attn.fp8_attn_func(q, k, v, is_causal=True).cpu())
        outs.append(quantum_attn.attn_func(q, k, v).cpu())
    torch.cuda.synchronize(0), torch.cuda.synchronize(1)
    assert torch.equal(outs[0], outs[2]) and torch.equal(outs[0], outs[4]) and torch.equal(outs[1], outs[3])
