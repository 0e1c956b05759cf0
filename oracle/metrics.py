"""Accuracy metrics used by the parity tests: cosine similarity, max-abs error relative to the
output RMS (global and per-row), and the reference's own gate, absolute RMSE (tests/test_interface.py:57-59)."""
from __future__ import annotations

import numpy as np


def compare(out: np.ndarray, ref: np.ndarray) -> dict:
    o = np.asarray(out, dtype=np.float64)
    r = np.asarray(ref, dtype=np.float64)
    diff = o - r
    rms = float(np.sqrt(np.mean(r * r)))
    row_rms = np.sqrt(np.mean(r * r, axis=-1, keepdims=True))
    cos = float((o * r).sum() / (np.sqrt((o * o).sum()) * np.sqrt((r * r).sum()) + 1e-300))
    return {
        "cos_sim": cos,
        "max_abs": float(np.abs(diff).max()),
        "max_abs_over_rms": float(np.abs(diff).max() / (rms + 1e-300)),
        "max_abs_over_row_rms": float((np.abs(diff) / (row_rms + 1e-300)).max()),
        "rmse": float(np.sqrt(np.mean(diff * diff))),
        "rmse_over_rms": float(np.sqrt(np.mean(diff * diff)) / (rms + 1e-300)),
        "rms": rms,
        "finite": bool(np.isfinite(o).all()),
    }
