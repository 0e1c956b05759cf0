"""Bit-level e4m3fn (1-4-3, bias 7, no inf, NaN = 0x7f/0xff, max 448) encode / decode in numpy.

Restates what ``tensor.to(torch.float8_e4m3fn)`` does to an fp32 value that is already inside [-448, 448]
(reference call site: src/quantum_attn/nn.py:18): round-to-nearest-even on the 3-bit mantissa, gradual underflow to
the subnormals m * 2^-9, and - because the reference clamps first - saturation never actually triggers, but the
encoder saturates anyway so it can also model ``cvt.rn.satfinite.e4m3x2.f32``.
"""
from __future__ import annotations

import numpy as np

E4M3_MAX = 448.0


def decode_table() -> np.ndarray:
    """float32[256]: value of every e4m3fn byte (NaN for 0x7f / 0xff)."""
    b = np.arange(256, dtype=np.int64)
    s = b >> 7
    e = (b >> 3) & 15
    m = b & 7
    v = np.where(e == 0, m * 2.0**-9, (8 + m) * np.exp2((e - 10).astype(np.float64)))
    v = np.where((e == 15) & (m == 7), np.nan, v)
    return np.where(s == 1, -v, v).astype(np.float32)


_TABLE = decode_table()


def e4m3_decode(b: np.ndarray) -> np.ndarray:
    return _TABLE[np.asarray(b, dtype=np.uint8)]


def e4m3_encode_rne_sat(x: np.ndarray) -> np.ndarray:
    """fp32 -> e4m3fn byte, RNE, saturating to +-448 (NaN -> 0x7f)."""
    x = np.asarray(x, dtype=np.float32)
    bits = x.view(np.uint32).astype(np.int64)
    sign = (bits >> 31) & 1
    a = np.abs(x.astype(np.float64))
    a = np.minimum(a, E4M3_MAX)
    # exponent of the quantisation step: normals have 3 mantissa bits, subnormals share step 2^-9
    with np.errstate(divide="ignore"):
        e = np.floor(np.log2(np.where(a > 0, a, 1.0)))
    e = np.maximum(e, -6.0)
    step = np.exp2(e - 3.0)
    q = a / step  # exact in fp64 (power-of-two scaling of an fp32 value)
    r = np.rint(q)  # numpy rint = round half to even
    a_q = r * step
    a_q = np.minimum(a_q, E4M3_MAX)
    # re-derive fields from the rounded value (rounding may have bumped the exponent)
    with np.errstate(divide="ignore"):
        e2 = np.floor(np.log2(np.where(a_q > 0, a_q, 1.0)))
    e2 = np.maximum(e2, -6.0)
    is_sub = a_q < 2.0**-6
    mant = np.where(is_sub, a_q / 2.0**-9, a_q / np.exp2(e2 - 3.0) - 8.0)
    expf = np.where(is_sub, 0.0, e2 + 7.0)
    byte = (sign << 7) | (expf.astype(np.int64) << 3) | mant.astype(np.int64)
    byte = np.where(np.isnan(x), 0x7F, byte)
    return byte.astype(np.uint8)
