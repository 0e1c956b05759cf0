"""External bar: NVIDIA's CuTe-DSL Blackwell FMHA example shipped inside the flashinfer wheel
(a warp-specialised persistent tcgen05 kernel), run with Float8E4M3FN and Float16 inputs on the C2 and C3 shapes.
Not the reference and not product code: a comparator on the same box.  The example's own `run()` returns its
benchmarked execution time in microseconds; this wrapper calls it and prints TFLOP/s with the reference's FLOP count.
Writes gpurun_out/cutedsl_fmha_bar.json."""
import importlib.util, json, os, sys
import flashinfer
import cutlass
import torch

path = os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass", "examples", "python", "CuTeDSL", "blackwell", "fmha.py")
sys.path.insert(0, os.path.dirname(path)); sys.path.insert(0, os.path.dirname(os.path.dirname(path)))
spec = importlib.util.spec_from_file_location("cutedsl_fmha", path)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
res = {}
SHAPES = {"C2_flux": ((1, 4608, 24, 128), False), "C3_llama": ((1, 8192, 32, 128), True)}
only = [a for a in sys.argv[1:] if a in SHAPES]  # optional: the workloads to run (bench.py passes its own)
for name, (shape, causal) in SHAPES.items():
    if only and name not in only:
        continue
    B, S, H, D = shape
    fl = 4 * B * H * S * S * D // (2 if causal else 1)
    for dt_name, dt in (("Float8E4M3FN", cutlass.Float8E4M3FN), ("Float16", cutlass.Float16)):
        try:
            torch.manual_seed(1111)
            us = mod.run(shape, shape, dt, cutlass.Float16, cutlass.Float32, cutlass.Float32, (128, 128), True, causal, False,
                         False, (-1, -1), 1.0, 1.0, 1.0, 1.0, 0.0, 0.1, 5, 20, True, False)
            res[f"{name}_{dt_name}"] = {"us": us, "tflops": fl / (us * 1e-6) / 1e12}
        except Exception as e:
            res[f"{name}_{dt_name}"] = {"error": repr(e)[:200]}
        print(name, dt_name, res[f"{name}_{dt_name}"], flush=True)
print("CUTEDSL_JSON " + json.dumps(res), flush=True)
if not only:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/cutedsl_fmha_bar.json", "w"), indent=1)
