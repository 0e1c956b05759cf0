"""BASELINE.json config 5 - the sweep D in {64, 128, 256} x S 1k ... 128k x causal - under parity, not only under the
stopwatch.  The reference gates its own sweep on correctness first (tests/test_interface.py:62-84 runs every (D, S,
causal) through `_test_attn_func` before timing it); the short end of the sweep is covered at full size by
tests/test_attention_gpu.py::test_parity_sweep, this file covers the LONG end, where the whole oracle would take
minutes: the CUDA path runs the full problem, the fp64 oracle scores a slice of it - 2 heads x 192 query rows (first
/ middle / last 64) against ALL keys - and size-independent properties cover the rest (finite everywhere, softmax
rows summing to one through V = 1).

Bounds: cosine similarity >= 0.999 in every mode; max-abs error <= 2e-2 of the row RMS for the modes built to meet it
("16bit": the reference's numerics and the library default; "fp8_hilo").
"""
import math

import pytest
import torch

import oracle
import quantum_attn
from quantumattention_b200 import _native

pytestmark = pytest.mark.gpu

ROW_BOUND = {"fp8": 0.30, "fp8_hilo": 0.02, "16bit": 0.02}


def _slice_oracle(q8, k8, sq, sk, vh, rows, causal, D):
    """fp64 softmax(Q K^T / sqrt(D)) V on the dequantised tensors for the chosen query rows (CPU)."""
    deq = lambda x8, sc: torch.from_numpy(oracle.dequantize(x8.view(torch.uint8).cpu().numpy(), sc.cpu().numpy())).double()
    qh, kh = deq(q8, sq)[:, :, rows], deq(k8, sk)
    sc = (qh @ kh.transpose(-1, -2)) / math.sqrt(D)
    if causal:
        S = kh.shape[2]
        sc = sc.masked_fill(torch.arange(S)[None, :] > rows[:, None], float("-inf"))
    return torch.softmax(sc, -1) @ vh


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("S", [32768, 131072])
@pytest.mark.parametrize("D", [64, 128, 256])
def test_long_sequences_oracle_on_2_heads_x_192_rows(D, S, causal):
    H = 2
    g = torch.Generator(device="cuda").manual_seed(D + S // 1024 + int(causal))
    q, k, v = (torch.randn((1, H, S, D), device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3))
    rows = torch.cat([torch.arange(0, 64), torch.arange(S // 2 - 32, S // 2 + 32), torch.arange(S - 64, S)])
    (q8, k8), (sq, sk) = _native.quantize_fp8([q, k], _native.QA_SCALE_HEAD)
    (v8,), (sv,) = _native.quantize_fp8([v], _native.QA_SCALE_HEAD)
    v16 = v.double().cpu()
    vdq = torch.from_numpy(oracle.dequantize(v8.view(torch.uint8).cpu().numpy(), sv.cpu().numpy())).double()
    refs = {"16bit": _slice_oracle(q8, k8, sq, sk, v16, rows, causal, D)}
    refs["fp8_hilo"] = refs["fp8"] = _slice_oracle(q8, k8, sq, sk, vdq, rows, causal, D)
    modes = ("16bit", "fp8_hilo", "fp8") if S <= 32768 else ("16bit", "fp8_hilo")
    for pv in modes:
        with quantum_attn.config.patch({"attention.pv_mode": pv}):
            out = quantum_attn.fp8_attn_func(q, k, v, is_causal=causal)
            ones = quantum_attn.fp8_attn_func(q, k, torch.ones_like(v), is_causal=causal)
        assert out.shape == q.shape and bool(torch.isfinite(out).all()), pv
        # (single-e4m3 P without the tensor-core row sum - D = 256 - normalises by the sum of the UNROUNDED weights:
        # the rows then sum to one only to within the e4m3 rounding of the weights)
        assert (ones.float() - 1.0).abs().max().item() < (0.03 if pv == "fp8" else 0.01), pv
        m = oracle.compare(out[:, :, rows.cuda()].float().cpu().numpy(), refs[pv].numpy())
        assert m["cos_sim"] >= 0.999 and m["max_abs_over_row_rms"] <= ROW_BOUND[pv], (D, S, causal, pv, m)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("D", [64, 256])
def test_long_sequences_16bit_entry_point(D, causal):
    """`attn_func` (bf16 Q, K, V; the other exported path) at S = 32768: slice oracle on the unquantised inputs."""
    S, H = 32768, 2
    g = torch.Generator(device="cuda").manual_seed(D + int(causal))
    q, k, v = (torch.randn((1, H, S, D), device="cuda", dtype=torch.bfloat16, generator=g) for _ in range(3))
    rows = torch.cat([torch.arange(0, 64), torch.arange(S // 2 - 32, S // 2 + 32), torch.arange(S - 64, S)])
    out = quantum_attn.attn_func(q, k, v, is_causal=causal)
    sc = (q.double().cpu()[:, :, rows] @ k.double().cpu().transpose(-1, -2)) / math.sqrt(D)
    if causal:
        sc = sc.masked_fill(torch.arange(S)[None, :] > rows[:, None], float("-inf"))
    ref = torch.softmax(sc, -1) @ v.double().cpu()
    m = oracle.compare(out[:, :, rows.cuda()].float().cpu().numpy(), ref.numpy())
    assert bool(torch.isfinite(out).all()) and m["cos_sim"] >= 0.9999 and m["max_abs_over_row_rms"] <= 0.02, m
