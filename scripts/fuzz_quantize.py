"""Developer tool: fuzz the sm_100a quantiser against the oracle with fresh random tensors; print any differing element."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from quantumattention_b200 import _native

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 200
shapes = [(1, 2, 1000, 256), (1, 3, 999, 128), (2, 8, 512, 64)]
modes = {"head-wise": _native.QA_SCALE_HEAD, "token-wise": _native.QA_SCALE_TOKEN, "head-wise-2pass": _native.QA_SCALE_HEAD_TWO_PASS}
bad = 0
for it in range(n_iter):
    g = torch.Generator().manual_seed(1000 + it)
    shape = shapes[it % len(shapes)]
    dtype = (torch.float16, torch.bfloat16)[(it // 3) % 2]
    x = (torch.randn(shape, generator=g) * torch.exp(torch.randn(shape[:2] + (1, shape[3]), generator=g))).to(dtype)
    for mode, m in modes.items():
        (x8,), (scale,) = _native.quantize_fp8([x.cuda()], m)
        torch.cuda.synchronize()
        b, s = oracle.quantize_fp8(x.float().numpy(), mode.replace("-2pass", ""))
        got = x8.view(torch.uint8).cpu().numpy()
        idx = np.argwhere(got != b)
        sc_ok = np.array_equal(scale.cpu().numpy(), s)
        if len(idx) or not sc_ok:
            bad += 1
            print(f"iter {it} {mode} {dtype} {shape}: {len(idx)} bytes differ, scales equal: {sc_ok}")
            for i in idx[:5]:
                i = tuple(i)
                xv = np.float32(x.float().numpy()[i])
                sv = np.float32(s[i[:2]] if mode != "token-wise" else s[i[:3]])
                print("   idx", i, "x", float(xv).hex(), "scale", float(sv).hex(), "q", float(np.float32(xv / sv)).hex(),
                      "got", hex(got[i]), "want", hex(b[i]))
print("fuzz done; failing cases:", bad)
