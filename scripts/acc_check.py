"""Developer tool: accuracy of the fused kernel (fp8 P mode) against the fp64 oracle on a few shapes; prints oracle.compare."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from quantumattention_b200 import _native
for (B, H, S, D, causal, kind) in [(1, 2, 1024, 128, False, "randn"), (1, 2, 1024, 128, True, "randn"),
                                   (1, 2, 2048, 64, False, "randn"), (1, 2, 1000, 128, False, "outlier_channels")]:
    q, k, v = oracle.make_qkv(B, H, S, S, D, seed=3, kind=kind)
    (q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q.cuda(), k.cuda(), v.cuda()], _native.QA_SCALE_HEAD)
    out, lse = _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=causal, sm_scale=1 / math.sqrt(D),
                                    p_mode=0, out_dtype=torch.bfloat16, return_lse=True)
    torch.cuda.synchronize()
    ref = oracle.fp8_attention_ref(q8.view(torch.uint8).cpu().numpy(), k8.view(torch.uint8).cpu().numpy(),
                                   v8.view(torch.uint8).cpu().numpy(), sq.cpu().numpy(), sk.cpu().numpy(),
                                   scale_v=sv.cpu().numpy(), is_causal=causal)
    m = oracle.compare(out.float().cpu().numpy(), ref.numpy())
    print((B, H, S, D, causal, kind), {k_: (round(v_, 6) if isinstance(v_, float) else v_) for k_, v_ in m.items()})
