"""CPU tests: the oracle (plain-C restatement + torch restatement) against the golden
fixtures recorded from the reference's own ``ms_deform_attn_core_pytorch``
(tests/golden/make_golden.py).  These pin the checker before the CUDA path trusts it."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import msda_oracle
from oracle.torch_ref import msda_core_torch

CASES = golden_names()


def tol(dtype):
    # fp32: both sides accumulate 16..32 products in different orders; fp64: ~1e-13.
    return dict(rtol=2e-5, atol=2e-6) if dtype == np.float32 else dict(rtol=1e-10, atol=1e-12)


def test_fixture_inventory():
    assert {"ref_test_f32", "ref_test_f64", "model_d32", "model_d36_l8", "border_f64",
            "border_fwd_f64", "batch_ragged", "model_d32_f64"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_forward_matches_reference(name):
    g = load_golden(name)
    out = msda_oracle.msda_forward(g["value"], g["shapes"], g["loc"], g["attn"])
    np.testing.assert_allclose(out, g["out"], **tol(g["value"].dtype))


@pytest.mark.parametrize("name", [c for c in CASES if c != "border_fwd_f64"])
def test_c_oracle_backward_matches_reference(name):
    g = load_golden(name)
    gv, gl, ga = msda_oracle.msda_backward(g["value"], g["shapes"], g["loc"], g["attn"], g["grad_out"])
    t = tol(g["value"].dtype)
    np.testing.assert_allclose(gv, g["grad_value"], **t)
    np.testing.assert_allclose(ga, g["grad_attn"], **t)
    # grad_loc is scaled by W/H and is a difference of products: looser absolute floor
    np.testing.assert_allclose(gl, g["grad_loc"], rtol=t["rtol"] * 5, atol=t["atol"] * 50)


@pytest.mark.parametrize("name", CASES)
def test_torch_restatement_matches_reference(name):
    g = load_golden(name)
    value, loc, attn = (torch.from_numpy(g[k]).requires_grad_(True) for k in ("value", "loc", "attn"))
    out = msda_core_torch(value, torch.from_numpy(g["shapes"]), loc, attn)
    t = tol(g["value"].dtype)
    np.testing.assert_allclose(out.detach().numpy(), g["out"], **t)
    if name == "border_fwd_f64":
        return
    gv, gl, ga = torch.autograd.grad((out * torch.from_numpy(g["grad_out"])).sum(), (value, loc, attn))
    np.testing.assert_allclose(gv.numpy(), g["grad_value"], **t)
    np.testing.assert_allclose(ga.numpy(), g["grad_attn"], **t)
    np.testing.assert_allclose(gl.numpy(), g["grad_loc"], rtol=t["rtol"] * 5, atol=t["atol"] * 50)


def test_c_oracle_out_of_range_samples_are_exact_zero():
    """Reference rule (.cuh:229, :359-362): a sample outside (-1,H)x(-1,W) contributes exactly 0
    to the output and to all three gradients."""
    shapes = np.array([[3, 5]], dtype=np.int64)
    value = np.random.default_rng(0).standard_normal((1, 15, 1, 4)).astype(np.float32)
    loc = np.full((1, 2, 1, 1, 1, 2), 7.0, dtype=np.float32)
    attn = np.ones((1, 2, 1, 1, 1), dtype=np.float32)
    assert not msda_oracle.msda_forward(value, shapes, loc, attn).any()
    gv, gl, ga = msda_oracle.msda_backward(value, shapes, loc, attn, np.ones((1, 2, 4), np.float32))
    assert not gv.any() and not gl.any() and not ga.any()


def test_c_oracle_linearity_in_value_and_attn():
    rng = np.random.default_rng(1)
    shapes = np.array([[5, 7], [3, 4]], dtype=np.int64)
    S = 35 + 12
    v1, v2 = (rng.standard_normal((2, S, 2, 8)) for _ in range(2))
    loc = rng.uniform(-0.1, 1.1, (2, 6, 2, 2, 3, 2))
    attn = rng.uniform(0, 1, (2, 6, 2, 2, 3))
    f = lambda v, a: msda_oracle.msda_forward(v, shapes, loc, a)
    np.testing.assert_allclose(f(2 * v1 - 3 * v2, attn), 2 * f(v1, attn) - 3 * f(v2, attn), rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(f(v1, 0.5 * attn), 0.5 * f(v1, attn), rtol=1e-12, atol=1e-13)
