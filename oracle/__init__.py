"""CPU oracle for the FP8 attention hot path (TEST INFRASTRUCTURE ONLY).

Nothing in ``quantumattention_b200`` imports this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker or the CPU baseline.

Parity status: the quantiser restatement is pinned bit-for-bit against the reference's own
``quantum_attn.nn._dynamically_quantize_fp8`` (reference: src/quantum_attn/nn.py:14-19) executed in this container on
fp32 inputs, and the attention restatement against ``quantum_attn.ops._fp8_attention_forward``
(reference: src/quantum_attn/ops.py:64-95); the vectors live in ``tests/golden/`` and were produced by
``oracle/gen_golden.py``.  The reference ships no golden vectors of its own.
"""
from .e4m3 import E4M3_MAX, decode_table, e4m3_decode, e4m3_encode_rne_sat  # noqa: F401
from .quantize_ref import dequantize, quantize_fp8  # noqa: F401
from .attention_ref import attention_flops, cpu_reference_step, fp8_attention_ref, sdpa_ref  # noqa: F401
from .metrics import compare  # noqa: F401
from .inputs import CONFIGS, make_qkv  # noqa: F401
from .ring_ref import attention_block_ref, head_scales, merge_ref, quantize_with_scale  # noqa: F401
