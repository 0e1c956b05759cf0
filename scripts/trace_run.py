"""Developer tool: run one C2 attention call on a -DQA_TRACE build and print the per-step clock stamps of CTA 0."""
import ctypes, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from quantumattention_b200 import _native
lib = _native.load(build_if_missing=False)
B, H, S, D = 1, 24, 4608, 128
q, k, v = oracle.make_qkv(B, H, S, S, D, seed=0)
(q8, k8, v8), (sq, sk, sv) = _native.quantize_fp8([q.cuda(), k.cuda(), v.cuda()], _native.QA_SCALE_HEAD)
tr = torch.zeros(4 * 80 * 8, dtype=torch.int64, device="cuda")
lib.qa_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.qa_debug_set_trace_cta(int(os.environ.get('TRACE_X', 0)), int(os.environ.get('TRACE_Y', 0)))
for it in range(3):
    tr.zero_()
    lib.qa_debug_set_trace(tr.data_ptr())
    _native.fp8_attn_fwd(q8, k8, v8, sq, sk, sv, scale_mode=0, is_causal=False, sm_scale=1 / math.sqrt(D), p_mode=0, out_dtype=torch.bfloat16)
    torch.cuda.synchronize()
t = tr.cpu().view(4, 80, 8)
t0 = int(t[t > 1000].min())
print('t0', t0)
def rel(x): return int(x) - t0 if int(x) > 1000 else -1
print("step | sm0: top ldwait maxdone arrive end nr | sm1: ... | mma0: waitP gotP issued | mma1")
for j in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    row = []
    for r in (0, 1):
        row.append(" ".join(f"{rel(t[r, j, e]):6d}" for e in range(5)) + f" {rel(t[r, j, 2]) - rel(t[r, j, 5])}")
    for r in (2, 3):
        row.append(" ".join(f"{rel(t[r, j, e]):6d}" for e in range(3)))
    print(f"{j:3d} | " + " | ".join(row))
for r, name in ((0, "sm0"), (1, "sm1")):
    d = t[r, 8:60]
    print(name, "mean step", float((d[1:, 0] - d[:-1, 0]).float().mean()), "ldwait", float((d[:, 1] - d[:, 0]).float().mean()),
          "max", float((d[:, 2] - d[:, 1]).float().mean()), "exp", float((d[:, 3] - d[:, 2]).float().mean()),
          "post", float((d[:, 4] - d[:, 3]).float().mean()),
          "wait for S_{j+1}: even steps", float((d[0::2, 2] - d[0::2, 5]).float().mean()), "odd steps", float((d[1::2, 2] - d[1::2, 5]).float().mean()))
for r, name in ((2, "mma0"), (3, "mma1")):
    d = t[r, 8:60]
    print(name, "waitP", float((d[:, 1] - d[:, 0]).float().mean()), "issue", float((d[:, 2] - d[:, 1]).float().mean()))
for r in (0, 1):
    e = [rel(t[r, 78, i]) for i in range(7)]
    print(f"tile{r} lifecycle: entry {e[0]} setup_done {e[1]} S0_ready {e[2]} loop_end {e[3]} o_full {e[4]} stored {e[5]} exit {e[6]}")
