"""Developer tool: what the host link gives - H2D alone (1 and 2 streams), D2H alone, and both directions at once."""
import torch, time
dev = torch.device("cuda:0")
N = 14155776  # elements of one C2 tensor (bf16)
h = [torch.empty(N, dtype=torch.bfloat16).pin_memory() for _ in range(3)]
d = [torch.empty(N, dtype=torch.bfloat16, device=dev) for _ in range(3)]
ho = torch.empty(N, dtype=torch.bfloat16).pin_memory()
do = torch.empty(N, dtype=torch.bfloat16, device=dev)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def up1():
    with torch.cuda.stream(s1):
        for a, b in zip(d, h): a.copy_(b, non_blocking=True)
def up2():
    with torch.cuda.stream(s1):
        d[0].copy_(h[0], non_blocking=True); d[1][:N // 2].copy_(h[1][:N // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        d[1][N // 2:].copy_(h[1][N // 2:], non_blocking=True); d[2].copy_(h[2], non_blocking=True)
def down():
    with torch.cuda.stream(s3): ho.copy_(do, non_blocking=True)
def both():
    up1(); down()
def both2():
    up2(); down()
B = 3 * N * 2
for name, fn, byts in (("H2D 1 stream", up1, B), ("H2D 2 streams", up2, B), ("D2H", down, N * 2), ("H2D + D2H (H2D GB/s)", both, B), ("H2D x2 + D2H (H2D GB/s)", both2, B)):
    t = run(fn)
    print(f"{name:28s} {t * 1e3:7.3f} ms  {byts / t / 1e9:6.1f} GB/s")
