"""CPU: pin both oracle restatements against the golden vectors produced by the reference's
own ms_deform_attn_core_pytorch (tests/golden/gen_golden.py), and against each other."""
import pytest
import torch

from conftest import golden_op_cases, load_golden
from oracle import msda_oracle as O

CASES = golden_op_cases()


def test_golden_present():
    assert len(CASES) >= 12
    assert "testpy_D32" in CASES and "d32_l4p4_oob" in CASES


@pytest.mark.parametrize("name", CASES)
def test_grid_sample_restatement_matches_reference_golden(name):
    g = load_golden("op", name)
    out, gv, gl, ga = O.core_grid_sample_fwd_bwd(g["value"], g["shapes"], g["loc"], g["attn"], g["grad_out"])
    # same torch build, same arithmetic path -> bit-for-bit
    assert torch.equal(out, g["out64"])
    assert torch.equal(gv, g["grad_value"])
    assert torch.equal(gl, g["grad_loc"])
    assert torch.equal(ga, g["grad_attn"])
    out32 = O.core_grid_sample(g["value"].float(), g["shapes"], g["loc"].float(), g["attn"].float())
    assert torch.equal(out32, g["out32"])


@pytest.mark.parametrize("name", CASES)
def test_c_loops_f64_match_reference_golden(name):
    g = load_golden("op", name)
    args = (g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"])
    out = O.forward_c(*args, precision="f64")
    gv, gl, ga = O.backward_c(*args, g["grad_out"], precision="f64")
    # different summation order / coordinate formula than grid_sample: fp64 round-off only
    assert O.rel_l2(out, g["out64"]) < 1e-13
    assert O.rel_l2(gv, g["grad_value"]) < 1e-13
    assert O.rel_l2(ga, g["grad_attn"]) < 1e-13
    assert O.rel_l2(gl, g["grad_loc"]) < 1e-12
    torch.testing.assert_close(out, g["out64"], rtol=1e-10, atol=1e-14)
    torch.testing.assert_close(gl, g["grad_loc"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_c_loops_f32_within_fp32_noise(name):
    g = load_golden("op", name)
    args = (g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"])
    out = O.forward_c(*args, precision="f32")
    gv, gl, ga = O.backward_c(*args, g["grad_out"], precision="f32")
    assert out.dtype == torch.float32
    # the fp32 inputs are roundings of the fp64 ones, so this is input + arithmetic noise
    assert O.rel_l2(out, g["out64"]) < 5e-6
    assert O.rel_l2(gv, g["grad_value"]) < 5e-6
    assert O.rel_l2(ga, g["grad_attn"]) < 5e-6
    mask = ~O.pixel_boundary_mask(g["loc"], g["shapes"], eps_px=1e-4)
    assert O.rel_l2(gl[mask], g["grad_loc"][mask]) < 5e-5


def test_out_of_range_samples_contribute_nothing():
    g = load_golden("op", "d32_l4p4_oob")
    loc = g["loc"].clone()
    far = (loc[..., 0] < -0.2) | (loc[..., 0] > 1.2) | (loc[..., 1] < -0.7) | (loc[..., 1] > 1.7)
    assert far.any()
    args = (g["value"], g["shapes"], g["level_start_index"])
    base = O.forward_c(*args, loc, g["attn"])
    attn2 = g["attn"].clone()
    attn2[far] = 123.0                                  # weights of far-outside samples are irrelevant
    assert torch.equal(O.forward_c(*args, loc, attn2), base)
    _, gl, ga = O.backward_c(*args, loc, g["attn"], g["grad_out"])
    assert (gl[far] == 0).all() and (ga[far] == 0).all()     # cuh:365-367


def test_linearity_and_adjoint_properties():
    """Size-independent properties later reused at full size on the GPU."""
    g = load_golden("op", "d32_l4p4_oob")
    sh, lsi, loc, attn = g["shapes"], g["level_start_index"], g["loc"], g["attn"]
    v1, v2 = g["value"], torch.randn_like(g["value"])
    f = lambda v, a=attn: O.forward_c(v, sh, lsi, loc, a)
    torch.testing.assert_close(f(2.5 * v1 - v2), 2.5 * f(v1) - f(v2), rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(f(v1, 3 * attn), 3 * f(v1), rtol=1e-12, atol=1e-12)
    # <f(v), g> == <v, grad_value(g)>  (the op is linear in value; grad_value is its adjoint)
    gv, _, ga = O.backward_c(v1, sh, lsi, loc, attn, g["grad_out"])
    lhs = (f(v2) * g["grad_out"]).sum()
    rhs = (v2 * gv).sum()
    assert abs(lhs - rhs) < 1e-10 * max(1.0, abs(lhs))
    # linear in attn as well: <f(v; a2), g> == <a2, grad_attn(g)>
    a2 = torch.randn_like(attn)
    lhs = (f(v1, a2) * g["grad_out"]).sum()
    assert abs(lhs - (a2 * ga).sum()) < 1e-10 * max(1.0, abs(lhs))


def test_constant_value_sums_weights_of_interior_samples():
    sh = torch.tensor([[7, 9], [3, 5]])
    lsi = torch.tensor([0, 63])
    gen = torch.Generator().manual_seed(5)
    loc = torch.rand(1, 11, 2, 2, 3, 2, generator=gen, dtype=torch.float64) * 0.6 + 0.2   # >=1px from borders
    attn = torch.rand(1, 11, 2, 2, 3, generator=gen, dtype=torch.float64)
    value = torch.ones(1, 78, 2, 4, dtype=torch.float64)
    out = O.forward_c(value, sh, lsi, loc, attn).view(1, 11, 2, 4)
    torch.testing.assert_close(out, attn.sum((-1, -2))[..., None].expand_as(out), rtol=1e-12, atol=1e-12)


def test_empty_queries():
    sh = torch.tensor([[2, 2]]); lsi = torch.tensor([0])
    value = torch.rand(1, 4, 1, 8, dtype=torch.float64)
    loc = torch.zeros(1, 0, 1, 1, 2, 2, dtype=torch.float64)
    attn = torch.zeros(1, 0, 1, 1, 2, dtype=torch.float64)
    assert O.forward_c(value, sh, lsi, loc, attn).shape == (1, 0, 8)
    assert O.core_grid_sample(value, sh, loc, attn).shape == (1, 0, 8)
