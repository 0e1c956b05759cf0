"""Developer tool: fused-kernel time for one shape (D S causal [BH]) with the library QA_NATIVE_LIB points at."""
import math, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
D, S, causal = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
BH = int(sys.argv[4]) if len(sys.argv) > 4 else max(1, (1 << 21) // S)
dev = torch.device("cuda:0")
q, k, v = (torch.randn((1, BH, S, D), device=dev, dtype=torch.bfloat16) for _ in range(3))
(q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q, k, v], _native.QA_SCALE_HEAD)
fl = 4.0 * BH * S * S * D / (2 if causal else 1)
_native.attn_events = []
for _ in range(12):
    _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D), p_mode=0, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
ev, _native.attn_events = _native.attn_events, None
ms = statistics.median(a.elapsed_time(b) for a, b in ev[2:])
print(f"{os.path.basename(os.environ.get('QA_NATIVE_LIB', 'default'))}: D={D} S={S} causal={int(causal)} BH={BH}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s")
