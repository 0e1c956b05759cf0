"""Developer tool: fused-kernel time over a list of shapes and P modes with the library QA_NATIVE_LIB points at.
usage: ab_kernels.py [tag]   (median of 15 CUDA-event-bracketed launches after 5 warm-ups, rotating 3 input sets)"""
import math, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native
tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("QA_NATIVE_LIB", "default"))
dev = torch.device("cuda:0")
PM = {"fp8": 0, "hilo": 1, "16bit": 2}
shapes = [("C2", 24, 4608, 128, False), ("C3", 32, 8192, 128, True), ("d64", 32, 8192, 64, False), ("d64c", 32, 8192, 64, True),
          ("d256", 16, 8192, 256, False), ("d256c", 16, 8192, 256, True), ("c4s", 4, 75600, 128, False),
          ("C2u", 24, 4606, 128, False)]  # C2u: S % 4 != 0 (token-wise: rows of scale_k not 16-byte aligned)
modes = os.environ.get("AB_MODES", "16bit,fp8,hilo").split(",")
only = os.environ.get("AB_SHAPES", "C2,C3,d64,d64c,d256,d256c,c4s")
if only:
    shapes = [s for s in shapes if s[0] in only.split(",")]
out = []
for name, H, S, D, causal in shapes:
    sets = []
    for i in range(3):
        q, k, v = (torch.randn((1, H, S, D), device=dev, dtype=torch.bfloat16) for _ in range(3))
        (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q, k, v], _native.QA_SCALE_HEAD)
        (q8t, k8t), (sqt, skt) = _native.quantize_fp8([q, k], _native.QA_SCALE_TOKEN) if "token" in modes else ((None, None), (None, None))
        sets.append((q8, k8, v, v8, sq, sk, sv, q8t, k8t, sqt, skt))
        del q, k
    fl = 4.0 * H * S * S * D / (2 if causal else 1)
    for mode in modes:
        def call(i):
            q8, k8, v, v8, sq, sk, sv, q8t, k8t, sqt, skt = sets[i % 3]
            if mode == "token":  # per-token scales, default P mode
                return _native.fp8_attn_fwd(q8t, k8t, v, sqt, skt, None, scale_mode=1, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                            p_mode=2, out_dtype=torch.bfloat16)
            if mode == "16bit":
                return _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                            p_mode=2, out_dtype=torch.bfloat16)
            return _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                        p_mode=PM[mode], out_dtype=torch.bfloat16)
        for i in range(5):
            call(i)
        torch.cuda.synchronize()
        _native.attn_events = []
        for i in range(15):
            call(i)
        torch.cuda.synchronize()
        ev, _native.attn_events = _native.attn_events, None
        ms = statistics.median(a.elapsed_time(b) for a, b in ev)
        out.append(f"{name}/{mode}: {ms * 1e3:.1f} us {fl / ms / 1e9:.0f} TF")
    del sets
print(f"[{tag}] " + " | ".join(out), flush=True)
