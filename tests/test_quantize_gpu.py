"""Parity of the sm_100a quantiser with the oracle: BIT-EXACT bytes and scales (integer / byte work)."""
import os

import numpy as np
import pytest
import torch

import oracle
import quantum_attn
from quantumattention_b200 import _native
from conftest import bf16_from_bits

pytestmark = pytest.mark.gpu

MODES = {"head-wise": _native.QA_SCALE_HEAD, "token-wise": _native.QA_SCALE_TOKEN}


def _check(x_cpu: torch.Tensor, mode: str):
    (x8,), (scale,) = _native.quantize_fp8([x_cpu.cuda()], MODES[mode])
    torch.cuda.synchronize()
    b, s = oracle.quantize_fp8(x_cpu.float().numpy(), mode)
    got = x8.view(torch.uint8).cpu().numpy()
    assert got.shape == b.shape
    nbad = int((got != b).sum())
    assert nbad == 0, f"{nbad} of {b.size} bytes differ"
    assert np.array_equal(scale.cpu().numpy(), s)


@pytest.mark.parametrize("mode,file", [("head-wise", "quantize_head.npz"), ("token-wise", "quantize_token.npz")])
def test_golden_vectors_from_reference(golden_dir, mode, file):
    g = np.load(os.path.join(golden_dir, file))
    x = bf16_from_bits(g["x_bf16_bits"])
    (x8,), (scale,) = _native.quantize_fp8([x.cuda()], MODES[mode])
    assert np.array_equal(x8.view(torch.uint8).cpu().numpy(), g["q_bytes"])
    assert np.array_equal(scale.cpu().numpy(), g["scale"])


@pytest.mark.parametrize("mode", ["head-wise", "token-wise"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 8, 512, 64), (1, 3, 999, 128), (1, 2, 1000, 256), (1, 1, 1, 64), (2, 2, 17, 128)])
def test_random_inputs(mode, dtype, shape):
    g = torch.Generator().manual_seed(hash((mode, str(dtype), shape)) % (2**31))
    x = (torch.randn(shape, generator=g) * torch.exp(torch.randn(shape[:2] + (1, shape[3]), generator=g))).to(dtype)
    _check(x, mode)


@pytest.mark.parametrize("kind", ["outlier_channels", "huge_token", "zero_head"])
@pytest.mark.parametrize("mode", ["head-wise", "token-wise"])
def test_stress_inputs(kind, mode):
    q, k, v = oracle.make_qkv(1, 3, 300, 300, 128, seed=5, kind=kind)
    for t in (q, k, v):
        _check(t, mode)


def test_three_tensors_one_launch_and_ragged_lengths():
    q, k, v = oracle.make_qkv(2, 4, 333, 470, 128, seed=9)
    (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q.cuda(), k.cuda(), v.cuda()], _native.QA_SCALE_HEAD)
    assert _native.last_launch_count() == 2
    for t, t8, s in ((q, q8, sq), (k, k8, sk), (v, v8, sv)):
        b, sc = oracle.quantize_fp8(t.float().numpy(), "head-wise")
        assert np.array_equal(t8.view(torch.uint8).cpu().numpy(), b)
        assert np.array_equal(s.cpu().numpy(), sc)


def test_strided_input_views():
    base = torch.randn(2, 300, 4, 128, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    x = base.permute(0, 2, 1, 3)  # [B,H,S,D] view of a [B,S,H,D] buffer, as DiT code often holds it
    assert not x.is_contiguous()
    (x8,), (scale,) = _native.quantize_fp8([x.cuda()], _native.QA_SCALE_HEAD)
    b, s = oracle.quantize_fp8(x.float().numpy(), "head-wise")
    assert np.array_equal(x8.view(torch.uint8).cpu().numpy(), b) and np.array_equal(scale.cpu().numpy(), s)


def test_public_dynamically_quantize_fp8():
    # reference signature: dynamically_quantize_fp8(t, *, reduction_dim=-1) -> (t_fp8, scale) (nn.py:22-42)
    x = torch.randn(2, 4, 100, 64, generator=torch.Generator().manual_seed(1)).to(torch.float16)
    t8, s = quantum_attn.dynamically_quantize_fp8(x.cuda())
    b, sc = oracle.quantize_fp8(x.float().numpy(), "token-wise")
    assert t8.dtype == torch.float8_e4m3fn and s.shape == (2, 4, 100) and s.dtype == torch.float32
    assert np.array_equal(t8.view(torch.uint8).cpu().numpy(), b) and np.array_equal(s.cpu().numpy(), sc)
    t8, s = quantum_attn.dynamically_quantize_fp8(x.cuda(), reduction_dim=[2, 3])
    b, sc = oracle.quantize_fp8(x.float().numpy(), "head-wise")
    assert s.shape == (2, 4)
    assert np.array_equal(t8.view(torch.uint8).cpu().numpy(), b) and np.array_equal(s.cpu().numpy(), sc)
    with pytest.raises(ValueError):
        quantum_attn.dynamically_quantize_fp8(x.cuda(), reduction_dim=0)


def test_full_size_checksum_flux_shape():
    # BASELINE config C2 (FLUX): too big for the numpy oracle in seconds -> size-independent properties instead:
    # (1) decode(encode(x)) is within half an e4m3 step of x / scale, (2) idempotence: re-quantising the dequantised
    # tensor reproduces the same bytes, (3) scale == amax/448 exactly.
    q, _, _ = oracle.make_qkv(1, 24, 4608, 4608, 128, seed=0)
    x = q.cuda()
    (x8,), (scale,) = _native.quantize_fp8([x], _native.QA_SCALE_HEAD)
    amax = x.float().abs().amax(dim=(2, 3))
    assert torch.equal(scale, (amax * (1.0 / 448.0)).clamp_min(torch.finfo(torch.float32).eps))
    deq = x8.float() * scale[:, :, None, None]
    y = x.float() / scale[:, :, None, None]
    err = (x8.float() - y).abs()
    step = torch.where(y.abs() < 2.0**-6, torch.tensor(2.0**-9, device="cuda"),
                       torch.exp2(torch.floor(torch.log2(y.abs().clamp_min(2.0**-6))) - 3))
    assert bool((err <= 0.5 * step * 1.0001).all())
    (x8b,), (scale_b,) = _native.quantize_fp8([deq.to(torch.bfloat16)], _native.QA_SCALE_HEAD)
    # bf16 rounding of deq may move a value across an e4m3 rounding boundary only if it was not representable;
    # e4m3 * power-of-two-ish scale is exactly representable in bf16 only when scale has <= 5 mantissa bits, so
    # compare on the fp32 path instead: quantise(deq) in torch with the same formula
    y2 = (deq / scale[:, :, None, None]).clamp(-448, 448).to(torch.float8_e4m3fn)
    assert torch.equal(y2.view(torch.uint8), x8.view(torch.uint8))
