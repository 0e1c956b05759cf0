"""Host-side mirror of the reference's interface: names, signatures, error conventions (no GPU needed)."""
import inspect

import pytest
import torch

import quantum_attn
from quantum_attn import quantum_attn_interface


def test_public_surface_matches_reference():
    # reference: src/quantum_attn/__init__.py:10-31
    assert sorted(quantum_attn.__all__) == sorted([
        "attn_func", "attn_func_with_fallback", "dynamically_quantize_fp8", "fp8_attn_func",
        "fp8_attn_func_with_fallback", "fp8_token_wise_attn_func", "fp8_token_wise_attn_func_with_fallback",
    ])
    for n in quantum_attn.__all__:
        assert callable(getattr(quantum_attn, n)) and callable(getattr(quantum_attn_interface, n))
    sig = inspect.signature(quantum_attn.fp8_attn_func)
    # the reference's parameters, in its order (src/quantum_attn/quantum_attn_interface.py:101-113); anything beyond them
    # must be an optional trailing keyword (scale_v: a value tensor quantised ahead of time)
    ref_params = ["query", "key", "value", "attn_mask", "dropout_p", "is_causal", "scale", "scale_q", "scale_k",
                  "scaling_method"]
    assert list(sig.parameters)[:len(ref_params)] == ref_params
    for extra in list(sig.parameters)[len(ref_params):]:
        assert sig.parameters[extra].kind is inspect.Parameter.KEYWORD_ONLY and sig.parameters[extra].default is None
    assert sig.parameters["scale"].kind is inspect.Parameter.KEYWORD_ONLY
    sig = inspect.signature(quantum_attn.fp8_token_wise_attn_func)
    assert "scaling_method" not in sig.parameters


def test_op_schemas_match_reference():
    s = str(torch.ops.quantum_attn.fp8_attention_forward.default._schema)
    assert s == ("quantum_attn::fp8_attention_forward(Tensor query, Tensor key, Tensor value, Tensor? scale_q=None, "
                 "Tensor? scale_k=None, Tensor? attn_mask=None, float dropout_p=0., bool is_causal=False, *, "
                 "float? scale=None) -> Tensor")
    assert "attn_func_with_fallback" in dir(torch.ops.quantum_attn)


def test_config_patch():
    assert quantum_attn.config.attention.force_eager_fallback is False
    with quantum_attn.config.patch({"attention.force_eager_fallback": True}):
        assert quantum_attn.config.attention.force_eager_fallback is True
    assert quantum_attn.config.attention.force_eager_fallback is False


def test_unsupported_inputs_raise_valueerror_with_reason():
    q = torch.randn(1, 2, 16, 64, dtype=torch.bfloat16)
    with pytest.raises(ValueError, match="CUDA device"):
        quantum_attn.fp8_attn_func(q, q, q)
    with pytest.raises(ValueError):
        quantum_attn.attn_func(q, q, q)
    with pytest.raises(ValueError):
        quantum_attn.dynamically_quantize_fp8(q)


def test_validator_reasons():
    from quantum_attn import nn

    q = torch.randn(1, 2, 16, 64, dtype=torch.bfloat16)
    v = nn._validate_input
    assert v(q, q, q, scaling_method="head-wise") == (False, "Expected query, key, and value to be on a CUDA device")
    assert v(q, q, q, attn_mask=q, scaling_method="head-wise")[1] == "NYI: attn_mask must be None"
    assert v(q, q, q, dropout_p=0.1, scaling_method="head-wise")[1] == "NYI: dropout_p must be 0.0"
    assert v(q, q, q, scaling_method="block")[1] == "Unsupported scaling_method: block"
    assert "same dtype" in v(q, q.half(), q, scaling_method="head-wise")[1]
    assert "value.dtype" in v(q, q, q.float(), scaling_method="head-wise")[1]
    qg = q.clone().requires_grad_()
    assert "leaf tensors" in v(qg, q, q, scaling_method="head-wise")[1]


def test_with_fallback_runs_sdpa_on_unsupported_input():
    q, k, v = (torch.randn(1, 2, 16, 64) for _ in range(3))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    for fn in (quantum_attn.attn_func_with_fallback, quantum_attn.fp8_attn_func_with_fallback,
               quantum_attn.fp8_token_wise_attn_func_with_fallback):
        assert torch.equal(fn(q, k, v, is_causal=True), ref)


def test_fake_tensor_shapes():
    from torch._subclasses.fake_tensor import FakeTensorMode

    with FakeTensorMode():
        q8 = torch.empty(2, 4, 100, 128, dtype=torch.float8_e4m3fn, device="cuda")
        k8 = torch.empty(2, 4, 77, 128, dtype=torch.float8_e4m3fn, device="cuda")
        v = torch.empty(2, 4, 77, 128, dtype=torch.float16, device="cuda")
        s = torch.empty(2, 4, device="cuda")
        out = torch.ops.quantum_attn.fp8_attention_forward(q8, k8, v, s, s)
        assert out.shape == (2, 4, 100, 128) and out.dtype == torch.float16
        t8, sc = quantum_attn.dynamically_quantize_fp8(v, reduction_dim=-1)
        assert t8.dtype == torch.float8_e4m3fn and sc.shape == (2, 4, 77)


def test_traced_graph_holds_the_native_ops_only():
    """dynamo trace (no GPU needed: the backend only looks at the graph): fp8_attn_func becomes ONE graph of the native
    quantiser op + the reference-named attention op - no aten arithmetic that Inductor would turn into its own
    kernels (the reference's graph holds its aten quantiser there, src/quantum_attn/nn.py:14-19,410-418)."""
    class Stop(Exception):
        pass

    seen = []

    def backend(gm, example_inputs):
        seen.append([str(n.target) for n in gm.graph.nodes if n.op == "call_function"])

        def run(*a):
            raise Stop()
        return run

    q = torch.randn(1, 2, 64, 64, dtype=torch.bfloat16)
    for fn, kw in ((quantum_attn.fp8_attn_func, {"is_causal": True}), (quantum_attn.fp8_token_wise_attn_func, {}),
                   (lambda a, b, c: quantum_attn.dynamically_quantize_fp8(a, reduction_dim=[2, 3]), {})):
        torch._dynamo.reset()
        seen.clear()
        with quantum_attn.config.patch({"attention.skip_supported_check": True}):
            with pytest.raises(Stop):
                torch.compile(fn, backend=backend, fullgraph=True)(q, q, q, **kw)
        assert len(seen) == 1
        names = seen[0]
        assert all(n.startswith("quantum_attn.") or "getitem" in n for n in names), names
        assert any("quantize" in n for n in names), names


def test_scale_shape_rules():
    from quantum_attn import nn, ops
    from quantumattention_b200 import _native

    q8 = torch.zeros(2, 3, 10, 64, dtype=torch.float8_e4m3fn)
    k8 = torch.zeros(2, 3, 12, 64, dtype=torch.float8_e4m3fn)
    assert ops._scale_mode_of(torch.ones(2, 3), q8, torch.ones(2, 3), k8) == _native.QA_SCALE_HEAD
    assert ops._scale_mode_of(torch.ones(2, 3, 10), q8, torch.ones(2, 3, 12), k8) == _native.QA_SCALE_TOKEN
    with pytest.raises(ValueError, match="granularity"):
        ops._scale_mode_of(torch.ones(2, 3), q8, torch.ones(2, 3, 12), k8)
    with pytest.raises(ValueError, match="matches neither"):
        ops._scale_mode_of(torch.ones(2, 3, 11), q8)
    v = nn._validate_input
    q = torch.randn(2, 3, 10, 64, dtype=torch.bfloat16)
    # a tensor is pre-quantised exactly when its scale comes with it; the key alone may be
    assert "scale_q and scale_k" in v(q, q, q, scaling_method="head-wise", scale_q=torch.ones(2, 3), scale_k=torch.ones(2, 3))[1]
    assert "scale_q and scale_k" in v(q8, k8, q, scaling_method="head-wise")[1]
    assert "CUDA device" in v(q, k8, q, scaling_method="head-wise", scale_k=torch.ones(2, 3))[1]  # passes the dtype rules
    assert "scale_v" in v(q, q, q8, scaling_method="head-wise")[1]
    assert "scale_v" in v(q, q, q, scaling_method="head-wise", scale_v=torch.ones(2, 3))[1]
