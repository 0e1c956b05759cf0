"""BASELINE.json configs[4]: sweep D in {64,128,256} x S in 1k..128k x causal in {0,1} on one GPU.

Per point: the fused FP8 attention kernel alone (CUDA events around each launch, inputs already quantised and far
larger than L2 in total), and the head-wise quantiser of Q, K, V (memset + kernel).  B*H is chosen so that every point
holds 2^21 tokens per tensor (the reference's own benchmark shape, B16 H16 S8192 - tests/test_interface.py:95-98 - is
the S = 8192 row).  Prints a markdown table (-> profiles/)."""
import math, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantumattention_b200 import _native

TOKENS = 1 << 21
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
name = torch.cuda.get_device_name(0)
print(f"# Sweep on {name}: fp8_attn_func pieces, head-wise scales, bf16 inputs, {TOKENS} tokens per tensor\n")
print("Fused kernel alone (CUDA events per launch, median) in the default mode `16bit` (FP8 QK^T, 16-bit P and V: the "
      "reference's numerics) and in the opt-in all-FP8 mode `fp8`; quantiser of Q, K, V (one call, events per call).\n")
print("| D | causal | S | B*H | 16bit: kernel us | TFLOP/s | fp8: kernel us | TFLOP/s | % of 4.5 PF | quantiser us (Q,K,V) | quantiser GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
smax = int(os.environ.get("SWEEP_SMAX", 131072))
for D in (64, 128, 256):
    for causal in (False, True):
        S = 1024
        while S <= smax:
            BH = max(1, TOKENS // S)
            g = torch.Generator(device=dev).manual_seed(S + D)
            q, k, v = (torch.randn((1, BH, S, D), device=dev, dtype=torch.bfloat16, generator=g) for _ in range(3))
            _native.quant_events = []
            for _ in range(4):
                (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q, k, v], _native.QA_SCALE_HEAD)
            torch.cuda.synchronize()
            qev, _native.quant_events = _native.quant_events, None
            q_ms = statistics.median(a.elapsed_time(b) for a, b in qev[1:])
            fl = 4.0 * BH * S * S * D / (2 if causal else 1)
            reps = max(3, min(30, int(0.15e15 / fl)))
            res = {}
            for mode in ("16bit", "fp8"):
                _native.attn_events = []
                for _ in range(reps + 2):
                    if mode == "fp8":
                        _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                             p_mode=0, out_dtype=torch.bfloat16)
                    else:
                        _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                             p_mode=2, out_dtype=torch.bfloat16)
                torch.cuda.synchronize()
                ev, _native.attn_events = _native.attn_events, None
                a_ms = statistics.median(a.elapsed_time(b) for a, b in ev[2:])
                res[mode] = (a_ms, fl / (a_ms * 1e-3) / 1e12)
            qbytes = 3 * (BH * S * D * 3 + 4 * BH)
            print(f"| {D} | {int(causal)} | {S} | {BH} | {res['16bit'][0] * 1e3:.1f} | {res['16bit'][1]:.0f} | {res['fp8'][0] * 1e3:.1f} | "
                  f"{res['fp8'][1]:.0f} | {100 * res['fp8'][1] / 4500:.1f} | {q_ms * 1e3:.1f} | {qbytes / (q_ms * 1e-3) / 1e9:.0f} |", flush=True)
            del q, k, v, q8, k8, v8
            S *= 2
