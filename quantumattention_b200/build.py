"""Build libqattn_sm100.so in-tree with nvcc (no torch headers: seconds, not minutes).

The reference JIT-builds its kernel through torch.utils.cpp_extension.load_inline on first use, hard-wired to sm_90a
(reference: src/quantum_attn/tk/attention.py:651-708).  Here the library is a plain C-ABI shared object compiled
ahead of time for sm_100a only; the built file is git-ignored but travels with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libqattn_sm100.so")
PROBE_PATH = os.path.join(HERE, "qa_probe")
SOURCES = ["api.cu", "quantize.cu", "merge.cu", "attn_fwd.cu", "attn_fwd16.cu"]
HEADERS = ["attn_fwd_kernel.cuh", "ptx.cuh", "tma_host.h", "qattn_internal.h", os.path.join("..", "..", "include", "qattn.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-std=c++17", "-O3", "-lineinfo",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libqattn_sm100.so must be built where the CUDA toolkit is installed")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build the library if it is missing or older than its sources.  Safe to call from several processes at once
    (one rank per GPU on a fresh tree): an inter-process lock lets one of them build, the others wait and re-check,
    and the library appears under its final name by an atomic rename, never half-written."""
    import fcntl

    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if not (force or _stale(LIB_PATH, deps)):
        return LIB_PATH
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale(LIB_PATH, deps):  # (another process may have built it while this one waited)
                tmp = LIB_PATH + f".tmp.{os.getpid()}"
                compile_and_link(tmp, srcs, extra=["-Xptxas=-v"] if verbose else [], verbose=verbose,
                                 objdir=os.path.join(HERE, "build", os.path.basename(LIB_PATH) + ".d"))
                os.replace(tmp, LIB_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


def compile_and_link(out: str, srcs, extra=(), verbose: bool = False, objdir: str = None) -> None:
    """One nvcc -c per translation unit, all at once (attn_fwd.cu dominates), then a link step."""
    from concurrent.futures import ThreadPoolExecutor

    objdir = objdir or os.path.join(HERE, "build", os.path.basename(out) + ".d")
    os.makedirs(objdir, exist_ok=True)
    objs = [os.path.join(objdir, os.path.basename(s) + ".o") for s in srcs]

    def one(pair):
        src, obj = pair
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        list(ex.map(one, zip(srcs, objs)))
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", out, *objs], check=True)


def build_probe(force: bool = False) -> str:
    src = os.path.join(CSRC, "probe.cu")
    deps = [src, os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "tma_host.h")]
    if force or _stale(PROBE_PATH, deps):
        subprocess.run([_nvcc(), *NVCC_FLAGS, "-o", PROBE_PATH, src], check=True)
    return PROBE_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--probe" in sys.argv:
        print(build_probe(force="--force" in sys.argv))
