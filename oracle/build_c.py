"""Compile the C restatement (oracle/quantize_ref.c) into oracle/_build/libqa_oracle.so and bind it with ctypes."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "quantize_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libqa_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def load():
    lib = ctypes.CDLL(build())
    lib.qa_oracle_quantize.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long]
    lib.qa_oracle_quantize.restype = None
    lib.qa_oracle_e4m3_encode.argtypes = [ctypes.c_float]
    lib.qa_oracle_e4m3_encode.restype = ctypes.c_uint8
    return lib


def quantize_c(x: np.ndarray, mode: str):
    """x float32 [..., S, D] -> (bytes, scale) using the C oracle."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    *lead, S, D = x.shape
    if mode == "head-wise":
        groups, length, sshape = int(np.prod(lead, dtype=np.int64)), S * D, tuple(lead)
    elif mode == "token-wise":
        groups, length, sshape = int(np.prod(lead, dtype=np.int64)) * S, D, tuple(lead) + (S,)
    else:
        raise ValueError(mode)
    out = np.empty(x.shape, dtype=np.uint8)
    scale = np.empty(groups, dtype=np.float32)
    load().qa_oracle_quantize(x.ctypes.data, out.ctypes.data, scale.ctypes.data, groups, length)
    return out, scale.reshape(sshape)


if __name__ == "__main__":
    print(build(force=True))
