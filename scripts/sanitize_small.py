"""Developer tool: a few small, ragged calls of every kernel family - meant to run under compute-sanitizer."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle, quantum_attn
from quantumattention_b200 import _native
for (B, H, S, D, causal) in [(1, 2, 333, 128, False), (1, 2, 200, 64, True), (1, 1, 130, 256, True)]:
    q, k, v = oracle.make_qkv(B, H, S, S, D, seed=1)
    qc, kc, vc = q.cuda(), k.cuda(), v.cuda()
    for pv in ("fp8", "fp8_hilo", "16bit"):
        with quantum_attn.config.patch({"attention.pv_mode": pv}):
            o = quantum_attn.fp8_attn_func(qc, kc, vc, is_causal=causal)
    o = quantum_attn.fp8_token_wise_attn_func(qc, kc, vc, is_causal=causal)
    o = quantum_attn.attn_func(qc, kc, vc, is_causal=causal)
    (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([qc, kc, vc], _native.QA_SCALE_HEAD_TWO_PASS)
    o, lse = _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                  p_mode=0, out_dtype=torch.bfloat16, return_lse=True)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
# token-wise scales through the shared-memory ring (aligned rows: bulk copies; 333 above: the producer lane's plain loads),
# several trips round the ring
q, k, v = (t.cuda() for t in oracle.make_qkv(1, 2, 1156, 1156, 128, seed=2))
o = quantum_attn.fp8_token_wise_attn_func(q, k, v)
# gated launch: flags set on another stream
(q8, k8), (sq, sk) = _native.quantize_fp8([q, k], _native.QA_SCALE_HEAD)
flags = torch.zeros(2, dtype=torch.int32, device="cuda")
side = torch.cuda.Stream()
torch.cuda.synchronize()
with torch.cuda.stream(side):
    torch.cuda._sleep(2_000_000)
    _native.set_flag(flags, 0, side.cuda_stream)
    _native.set_flag(flags, 1, side.cuda_stream)
o2 = _native.fp8_attn_fwd(q8, k8, v, sq, sk, None, scale_mode=0, is_causal=False, sm_scale=1 / math.sqrt(128), p_mode=2,
                          out_dtype=torch.bfloat16, gate=(flags, 1, 1))
torch.cuda.synchronize()
assert torch.isfinite(o.float()).all() and torch.isfinite(o2.float()).all()
print("sanitize_small ok")
