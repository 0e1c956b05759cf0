"""Developer tool: host-side cost of one fp8_attn_func call (python + ctypes + allocator), GPU not waited for."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import quantum_attn
q, k, v = (torch.randn((1, 24, 4608, 128), device="cuda", dtype=torch.bfloat16) for _ in range(3))
for _ in range(20):
    quantum_attn.fp8_attn_func(q, k, v)
torch.cuda.synchronize()
for name, fn in (("fp8_attn_func", lambda: quantum_attn.fp8_attn_func(q, k, v)), ("attn_func", lambda: quantum_attn.attn_func(q, k, v))):
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(40):
            fn()
        best = min(best, (time.perf_counter() - t0) / 40 * 1e6)
        torch.cuda.synchronize()
    print(f"{name}: host {best:.1f} us per call")
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(40):
    quantum_attn.fp8_attn_func(q, k, v)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
