"""quantumattention_b200 - a from-scratch, B200-native (sm_100a) implementation of the FP8 fused-attention hot path
of WaveSpeedAI/QuantumAttention, behind the reference's own Python entry points
(reference: src/quantum_attn/__init__.py:10-31).  ``import quantum_attn`` (the alias package at the repo root)
resolves to this package, so reference users switch without touching call sites.
"""
from . import config, nn, ops, quantum_attn_interface  # noqa: F401
from .nn import QuantizedKV, quantize_kv  # noqa: F401  (extension: quantise K / V once, attend many times)
from .quantum_attn_interface import (
    attn_func,
    attn_func_with_fallback,
    dynamically_quantize_fp8,
    fp8_attn_func,
    fp8_attn_func_with_fallback,
    fp8_token_wise_attn_func,
    fp8_token_wise_attn_func_with_fallback,
)

__version__ = "0.1.0"

__all__ = [
    "attn_func",
    "attn_func_with_fallback",
    "dynamically_quantize_fp8",
    "fp8_attn_func",
    "fp8_attn_func_with_fallback",
    "fp8_token_wise_attn_func",
    "fp8_token_wise_attn_func_with_fallback",
]
