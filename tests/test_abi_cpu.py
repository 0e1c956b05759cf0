"""The C-ABI library must load without a GPU and export every symbol include/qattn.h declares."""
import ctypes
import os
import re

import pytest

from quantumattention_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "qattn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qa_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build_library()
    assert os.path.exists(path)
    lib = _native.load()
    assert lib.qa_abi_version() == _native.ABI_VERSION


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(build.build_library())
    declared = _declared_functions()
    assert set(declared) == set(_native.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_header_is_plain_c_and_the_c_example_compiles(tmp_path):
    """The boundary is a C ABI: include/qattn.h must compile as C99 (no C++-isms, no CUDA / torch types), and the C
    caller in examples/ must compile against it and link against the built library."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = os.path.join(ROOT, "examples", "c_abi_example.c")
    obj = str(tmp_path / "ex.o")
    subprocess.run([gcc, "-std=c99", "-fPIC", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj],
                   check=True)
    # every qa_ symbol the example uses resolves in the library
    lib = build.build_library()
    so = str(tmp_path / "libex.so")
    subprocess.run([gcc, "-shared", "-o", so, obj, lib, "-Wl,--no-undefined"], check=True)


def test_argument_validation_without_gpu():
    lib = _native.load()
    vp = ctypes.c_void_p
    one = (vp * 1)(16)
    strides = (ctypes.c_int64 * 4)(1, 1, 1, 1)
    S = (ctypes.c_int * 1)(8)
    # bad head dim -> QA_ERR_INVALID with the reference's wording
    rc = lib.qa_quantize_fp8(1, one, 0, strides, one, one, None, 1, 1, S, 96, 1, None)
    assert rc == -1 and b"Unsupported head dimension: 96" in lib.qa_last_error()
    rc = lib.qa_quantize_fp8(4, one, 0, strides, one, one, None, 1, 1, S, 64, 1, None)
    assert rc == -1
    rc = lib.qa_fp8_attn_fwd(16, 16, 16, 2, None, None, None, 16, 16, 16, 0, 16, 0, None, 1, 3, 2, 8, 8, 64, 0, 0.125, 0, None)
    assert rc == -1 and b"multiple of Hkv" in lib.qa_last_error()
    rc = lib.qa_fp8_attn_fwd(16, 16, 16, 0, None, None, None, 16, 16, None, 0, 16, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 0, None)
    assert rc == -1  # fp8 P mode with a 16-bit V
    # gated launch: flags are required, and a sane gate geometry
    rc = lib.qa_fp8_attn_fwd_gated(16, 16, 16, 2, None, None, None, 16, 16, 16, 0, 32, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 2,
                                   None, 1, 1, None)
    assert rc == -1
    rc = lib.qa_fp8_attn_fwd_gated(16, 16, 16, 2, None, None, None, 16, 16, 16, 0, 32, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 2,
                                   16, 0, 1, None)
    assert rc == -1 and b"gated launch" in lib.qa_last_error()
    assert lib.qa_set_flag(None, None) == -1
    rc = lib.qa_attn_fwd(16, 16, 16, 2, None, None, None, 16, None, 1, 2, 2, 8, 8, 64, 0, 0.125, None)
    assert rc == -1 and b"fp16 or bf16" in lib.qa_last_error()  # e4m3 inputs belong to qa_fp8_attn_fwd
    rc = lib.qa_attn_fwd(16, 16, 16, 0, None, None, None, 16, None, 1, 2, 2, 8, 8, 96, 0, 0.125, None)
    assert rc == -1 and b"Unsupported head dimension: 96" in lib.qa_last_error()
    import torch

    if not torch.cuda.is_available():
        # valid arguments but no device: must fail loudly, never fall back
        # (fake pointers: 16-byte aligned inputs, a 32-byte aligned output as the header asks)
        rc = lib.qa_fp8_attn_fwd(16, 16, 16, 2, None, None, None, 16, 16, 16, 0, 32, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 0, None)
        assert rc in (-2, -3)
        rc = lib.qa_attn_fwd(16, 16, 16, 0, None, None, None, 32, None, 1, 2, 2, 8, 8, 64, 0, 0.125, None)
        assert rc in (-2, -3)
    rc = lib.qa_fp8_attn_fwd(16, 16, 16, 2, None, None, None, 16, 16, 16, 0, 16, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 0, None)
    assert rc == -1 and b"32-byte aligned" in lib.qa_last_error()

    # strides: every one a multiple of 16 bytes, the row stride at least the head dimension
    bad = (ctypes.c_int64 * 3)(8 * 64 * 2, 8 * 64, 72)
    rc = lib.qa_fp8_attn_fwd(16, 16, 16, 2, bad, None, None, 16, 16, 16, 0, 32, 0, None, 1, 2, 2, 8, 8, 64, 0, 0.125, 0,
                             None)
    assert rc == -1 and b"multiple of 16 bytes" in lib.qa_last_error()
    # the one-call entry point validates like the two it combines
    rc = lib.qa_fp8_attn_func(16, 16, 16, 0, None, None, None, 16, 16, None, 16, 16, None, 16, 0, 32, None,
                              1, 2, 2, 8, 8, 96, 0, 0.125, 0, 2, None)
    assert rc == -1 and b"Unsupported head dimension: 96" in lib.qa_last_error()
    rc = lib.qa_fp8_attn_func(16, 16, 16, 0, None, None, None, 16, 16, None, 16, 16, None, 16, 0, 32, None,
                              1, 2, 2, 8, 8, 64, 0, 0.125, 0, 0, None)
    assert rc == -1 and b"v8 and scale_v" in lib.qa_last_error()
