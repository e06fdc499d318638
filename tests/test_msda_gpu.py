"""GPU parity tests: the sm_100a kernels (through the C ABI, via MSDeformAttnFunction / torch.ops.msda)
against the oracle on identical seeded inputs.  Run on the B200 box:  pytest tests -m gpu

Tolerance contract (BASELINE.json north_star) -- kernel vs the fp64 oracle evaluated
on the SAME (already rounded) inputs:
  fp64 : rel-L2 <= 1e-12 everywhere (arithmetic order differs from grid_sample, nothing else)
  fp32 : forward rel-L2 <= 1e-5 and max-abs/max <= 2e-5;
         grad_value, grad_attn rel-L2 <= 1e-4 (atomic ordering);
         grad_loc rel-L2 <= 1e-4 after masking samples within 1e-4 px of a pixel boundary
         (d out / d loc is discontinuous there and floor() is precision dependent)
  bf16 : value/grad_out rounded to bf16, fp32 arithmetic: forward rel-L2 <= 4e-3 (one bf16 rounding
         of the output), grad_value rel-L2 <= 1e-2 (one bf16 rounding), grad_loc/grad_attn <= 1e-4
         (fp32 outputs, masked as above)
"""
import ctypes

import pytest
import torch

from conftest import golden_op_cases, load_golden
from oracle import msda_oracle as O

pytestmark = pytest.mark.gpu

TOL = {
    torch.float64: dict(fwd=1e-12, fwd_max=1e-11, gv=1e-12, ga=1e-12, gl=1e-11),
    torch.float32: dict(fwd=1e-5, fwd_max=2e-5, gv=1e-4, ga=1e-4, gl=1e-4),
    torch.bfloat16: dict(fwd=4e-3, fwd_max=2e-2, gv=1e-2, ga=1e-4, gl=1e-4),
}


@pytest.fixture(scope="module")
def msda(cuda_device):
    import monosowa_b200
    return monosowa_b200


def _levels(shapes):
    sh = torch.as_tensor(shapes, dtype=torch.long)
    return sh, torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))


def run_ours(msda, value, sh, lsi, loc, attn, grad_out, step=64):
    """forward + autograd backward through the reference-shaped API."""
    dev = torch.device("cuda:0")
    v = value.to(dev).requires_grad_(True)
    l = loc.to(dev).requires_grad_(True)
    a = attn.to(dev).requires_grad_(True)
    out = msda.MSDeformAttnFunction.apply(v, sh.to(dev), lsi.to(dev), l, a, step)
    out.backward(grad_out.to(dev).reshape(out.shape))
    torch.cuda.synchronize()
    return out.detach().cpu(), v.grad.cpu(), l.grad.cpu(), a.grad.cpu()


def check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, dtype, label=""):
    """Inputs are fp64 'master' tensors; they are rounded to the kernel dtypes first, and the fp64
    oracle is evaluated on those rounded values."""
    ct = torch.float64 if dtype == torch.float64 else torch.float32
    v_k, g_k = value.to(dtype), grad_out.to(dtype)
    l_k, a_k = loc.to(ct), attn.to(ct)
    out, gv, gl, ga = run_ours(msda, v_k, sh, lsi, l_k, a_k, g_k)
    assert out.dtype == dtype and gv.dtype == dtype and gl.dtype == ct and ga.dtype == ct
    args = (v_k.double(), sh, lsi, l_k.double(), a_k.double())
    ref_out = O.forward_c(*args)
    ref_gv, ref_gl, ref_ga = O.backward_c(*args, g_k.double())
    tol = TOL[dtype]
    e = dict(fwd=O.rel_l2(out, ref_out), fwd_max=O.max_abs_over_max(out, ref_out),
             gv=O.rel_l2(gv, ref_gv), ga=O.rel_l2(ga, ref_ga))
    keep = ~O.pixel_boundary_mask(l_k, sh, eps_px=1e-4)
    e["gl"] = O.rel_l2(gl[keep], ref_gl[keep])
    bad = {k: (v, tol[k]) for k, v in e.items() if not v <= tol[k]}
    assert not bad, f"{label} {dtype}: out of tolerance {bad}; all={e}; masked={int((~keep).sum()) // 2}"
    return e


# --------------------------------------------------------------------------------------------
# 1. golden vectors from the reference's own ms_deform_attn_core_pytorch
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_op_cases())
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_golden_vectors(msda, name, dtype):
    g = load_golden("op", name)
    ct = dtype
    out, gv, gl, ga = run_ours(msda, g["value"].to(dtype), g["shapes"], g["level_start_index"],
                               g["loc"].to(ct), g["attn"].to(ct), g["grad_out"].to(dtype))
    if dtype == torch.float64:
        assert O.rel_l2(out, g["out64"]) < 1e-12
        assert O.rel_l2(gv, g["grad_value"]) < 1e-12
        assert O.rel_l2(ga, g["grad_attn"]) < 1e-12
        assert O.rel_l2(gl, g["grad_loc"]) < 1e-11
    else:
        # inputs were rounded to fp32 -> compare with the golden at fp32 input-noise level
        assert O.rel_l2(out, g["out64"]) < 5e-6
        assert O.rel_l2(gv, g["grad_value"]) < 5e-6
        assert O.rel_l2(ga, g["grad_attn"]) < 5e-6
        keep = ~O.pixel_boundary_mask(g["loc"], g["shapes"], eps_px=1e-4)
        assert O.rel_l2(gl[keep], g["grad_loc"][keep]) < 5e-5


# --------------------------------------------------------------------------------------------
# 2. the reference's own test contract (ops/test.py)
# --------------------------------------------------------------------------------------------
def _testpy_inputs(D, dtype, dev):
    torch.manual_seed(3)                                                     # ops/test.py:28
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    sh, lsi = _levels([(6, 4), (3, 2)])
    S = 30
    value = (torch.rand(N, S, M, D) * 0.01).to(dev, dtype)
    loc = torch.rand(N, Lq, M, L, P, 2).to(dev, dtype)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).to(dev, dtype)
    return value, sh.to(dev), lsi.to(dev), loc, attn


def test_refcontract_forward_equal_with_pytorch_double(msda, cuda_device):
    """ops/test.py:31-44 -- default allclose against the grid_sample path, fp64."""
    value, sh, lsi, loc, attn = _testpy_inputs(2, torch.float64, cuda_device)
    ref = O.core_grid_sample(value.cpu(), sh.cpu(), loc.cpu(), attn.cpu())
    out = msda.MSDeformAttnFunction.apply(value, sh, lsi, loc, attn, 2).cpu()
    assert torch.allclose(out, ref)


def test_refcontract_forward_equal_with_pytorch_float(msda, cuda_device):
    """ops/test.py:47-60 -- rtol 1e-2 / atol 1e-3, fp32."""
    value, sh, lsi, loc, attn = _testpy_inputs(2, torch.float32, cuda_device)
    ref = O.core_grid_sample(value.cpu(), sh.cpu(), loc.cpu(), attn.cpu())
    out = msda.MSDeformAttnFunction.apply(value, sh, lsi, loc, attn, 2).cpu()
    assert torch.allclose(out, ref, rtol=1e-2, atol=1e-3)
    assert O.rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize("D", [30, 32, 64, 71])
def test_refcontract_gradcheck_all_inputs(msda, cuda_device, D):
    """ops/test.py:63-78 -- numerical gradient check in fp64 (all three differentiable inputs)."""
    value, sh, lsi, loc, attn = _testpy_inputs(D, torch.float64, cuda_device)
    value.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    assert torch.autograd.gradcheck(msda.MSDeformAttnFunction.apply, (value, sh, lsi, loc, attn, 2))


@pytest.mark.parametrize("D", [1025, 2048, 3096])
def test_refcontract_gradcheck_wide_channels(msda, cuda_device, D):
    """ops/test.py:85 lists D up to 3096 (one per reference kernel branch).  A full numerical
    Jacobian w.r.t. value is 6e4..2e5 forward calls, so: numerical check w.r.t. loc and attn (small),
    and grad_value against the analytic C oracle."""
    value, sh, lsi, loc, attn = _testpy_inputs(D, torch.float64, cuda_device)
    loc.requires_grad_(True); attn.requires_grad_(True)
    fn = lambda l, a: msda.MSDeformAttnFunction.apply(value, sh, lsi, l, a, 2)
    assert torch.autograd.gradcheck(fn, (loc, attn))
    g = torch.randn(1, 2, 2 * D, dtype=torch.float64)
    _, gv, gl, ga = run_ours(msda, value.detach().cpu(), sh.cpu(), lsi.cpu(), loc.detach().cpu(), attn.detach().cpu(), g)
    rgv, rgl, rga = O.backward_c(value.cpu(), sh.cpu(), lsi.cpu(), loc.detach().cpu(), attn.detach().cpu(), g)
    assert O.rel_l2(gv, rgv) < 1e-12 and O.rel_l2(gl, rgl) < 1e-11 and O.rel_l2(ga, rga) < 1e-12


# --------------------------------------------------------------------------------------------
# 3. vector kernels vs the explicit-loop oracle: dtypes, D, ragged tails, out-of-range samples
# --------------------------------------------------------------------------------------------
def _random_case(seed, shapes, N, M, D, Lq, P, spread=2.32, shift=-0.66):
    g = torch.Generator().manual_seed(seed)
    sh, lsi = _levels(shapes)
    S, L = int(sh.prod(1).sum()), len(shapes)
    value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64)
    loc = torch.rand(N, Lq, M, L, P, 2, generator=g, dtype=torch.float64) * spread + shift
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, dtype=torch.float64), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g, dtype=torch.float64)
    return value, sh, lsi, loc, attn, grad_out


VEC_CASES = [
    # (shapes, N, M, D, Lq, P)
    ([(12, 40), (6, 20), (3, 10), (2, 5)], 2, 8, 32, 131, 4),     # MonoDETR head layout, ragged Lq
    ([(12, 40), (6, 20), (3, 10), (2, 5)], 1, 8, 32, 1, 4),       # a single query
    ([(9, 7), (5, 4)], 3, 3, 32, 17, 4),                          # M not a power of two
    ([(9, 7), (5, 4), (1, 1)], 2, 2, 16, 33, 4),
    ([(9, 7), (5, 4)], 2, 4, 64, 19, 4),
    ([(16, 16)], 1, 1, 32, 257, 4),                               # single level, many chunks
    ([(9, 7), (5, 4)], 2, 4, 32, 19, 2),                          # L*P = 4 < lanes per group
    ([(9, 7), (5, 4), (3, 3)], 2, 4, 32, 19, 3),                  # L*P = 9: ragged last batch
    ([(9, 7), (5, 4), (3, 3)], 1, 2, 64, 5, 5),                   # L*P = 15 with 16-lane groups
    ([(9, 7), (5, 4)], 2, 4, 16, 21, 8),                          # L*P = 16 with 4-lane groups
    ([(9, 7), (5, 4)], 2, 4, 24, 19, 4),                          # D not vectorisable -> generic
]


@pytest.mark.parametrize("case", VEC_CASES, ids=lambda c: f"M{c[2]}D{c[3]}Lq{c[4]}P{c[5]}L{len(c[0])}")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
def test_kernels_vs_oracle(msda, case, dtype):
    shapes, N, M, D, Lq, P = case
    value, sh, lsi, loc, attn, grad_out = _random_case(100 + D + Lq, shapes, N, M, D, Lq, P)
    check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, dtype, label=str(case))


def _needs_ab(msda):
    if not msda._lib.has_ab_flavours():
        pytest.skip("tile kernels exist only in the measurement build: MSDA_AB=1 pytest ... (python -m monosowa_b200.build --ab)")


@pytest.mark.parametrize("variant", [(11, 11), (11, 21), (12, 20), (99, 99)], ids=["record", "binned", "tile", "generic"])
def test_every_launch_variant_is_correct(msda, variant):
    """record kernels, binned backward and tile kernels (forced for a short query set) and the generic kernels on a
    vectorisable shape"""
    if variant == (12, 20):
        _needs_ab(msda)
    value, sh, lsi, loc, attn, grad_out = _random_case(7, [(12, 40), (6, 20), (3, 10), (2, 5)], 2, 8, 32, 203, 4)
    L = msda._lib
    try:
        L.set_tuning("fwd_variant", variant[0]); L.set_tuning("bwd_variant", variant[1])
        check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, torch.float32, label=f"variant{variant}")
        check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, torch.bfloat16, label=f"variant{variant}")
    finally:
        L.set_tuning("fwd_variant", -1); L.set_tuning("bwd_variant", -1)


def _fuzz_case(i):
    g = torch.Generator().manual_seed(9000 + i)
    r = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
    L = r(1, 5)
    shapes = [(r(1, 9), r(1, 11)) for _ in range(L)]
    D = [16, 32, 64, 32, 32, 8, 24, 30, 5][r(0, 8)]
    dtype = [torch.float32, torch.float32, torch.bfloat16, torch.float64][r(0, 3)]
    return shapes, r(1, 3), [1, 2, 3, 4, 8][r(0, 4)], D, r(1, 70), r(1, 6), dtype


@pytest.mark.parametrize("i", range(32))
def test_fuzz_random_shapes(msda, i):
    """seeded random (levels, N, M, D, Lq, P, dtype): record kernels for D in {16,32,64} (any L*P, ragged
    batches, 1x1 levels), generic kernels otherwise; ~25 % of the locations outside [0,1]."""
    shapes, N, M, D, Lq, P, dtype = _fuzz_case(i)
    value, sh, lsi, loc, attn, grad_out = _random_case(9100 + i, shapes, N, M, D, Lq, P, spread=1.5, shift=-0.25)
    check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, dtype, label=f"fuzz{i} {shapes} N{N} M{M} D{D} Lq{Lq} P{P}")


def test_generic_and_record_kernels_agree(msda):
    """variant 99 forces the generic kernels for a vectorisable shape."""
    value, sh, lsi, loc, attn, grad_out = _random_case(8, [(12, 40), (6, 20), (3, 10), (2, 5)], 2, 8, 32, 77, 4)
    L = msda._lib
    a = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    try:
        L.set_tuning("fwd_variant", 99); L.set_tuning("bwd_variant", 99)
        b = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    finally:
        L.set_tuning("fwd_variant", -1); L.set_tuning("bwd_variant", -1)
    for x, y, tol in zip(a, b, (2e-6, 1e-5, 1e-5, 1e-5)):
        assert O.rel_l2(x, y) < tol


@pytest.mark.parametrize("D", [32, 64])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_forward_lane_widths_agree_bitwise(msda, cuda_device, D, dtype):
    """the forward gathers 8 channels per lane (one 256-bit / 128-bit load) when D >= 32 and `value` is aligned to
    8 elements, 4 channels per lane otherwise (fwd_pipe = 4 forces it): same samples in the same order per channel,
    so the outputs are bit-identical -- checked through the knob and through a `value` view that is 16- but not
    32-byte aligned."""
    value, sh, lsi, loc, attn, _ = _random_case(21, [(12, 40), (6, 20), (3, 10), (2, 5)], 2, 4, D, 333, 4, spread=1.4, shift=-0.2)
    dev = cuda_device
    v = value.to(dtype).to(dev)
    args = (sh.to(dev), lsi.to(dev), loc.float().to(dev), attn.float().to(dev))
    wide = torch.ops.msda.forward(v, *args, 64)
    ref = O.forward_c(v.double().cpu(), sh, lsi, loc.float().double(), attn.float().double())
    assert O.rel_l2(wide, ref) < TOL[dtype]["fwd"]
    msda._lib.set_tuning("fwd_pipe", 4)
    try:
        narrow = torch.ops.msda.forward(v, *args, 64)
    finally:
        msda._lib.set_tuning("fwd_pipe", -1)
    assert torch.equal(wide, narrow)
    shift = 16 // v.element_size()                       # elements in 16 bytes
    flat = torch.empty(v.numel() + shift, dtype=dtype, device=dev)
    off = shift if flat.data_ptr() % 32 == 0 else 0
    view = flat[off:off + v.numel()].view_as(v)
    view.copy_(v)
    assert view.data_ptr() % 32 == 16
    assert torch.equal(torch.ops.msda.forward(view, *args, 64), wide)
    # the record backward takes 8 channels per lane for D = 64 only (bwd_pipe = 4 forces 4): the per-lane partial
    # dot products are summed in a different order, so agreement is to rounding
    g = torch.randn_like(wide)
    b_wide = torch.ops.msda.backward(v, *args, g, 64)
    msda._lib.set_tuning("bwd_pipe", 4)
    try:
        b_narrow = torch.ops.msda.backward(v, *args, g, 64)
    finally:
        msda._lib.set_tuning("bwd_pipe", -1)
    b_view = torch.ops.msda.backward(view, *args, g, 64)
    for x, y, z, tol in zip(b_wide, b_narrow, b_view, (1e-2 if dtype == torch.bfloat16 else 1e-5, 1e-5, 1e-5)):
        assert O.rel_l2(x, y) < tol and O.rel_l2(z, y) < tol


# --- binned backward (msda_backward_binned.cu; the default for long query sets) and the tile kernels (msda_tiles.cuh,
# fwd_tile_kernel, msda_backward_tiled.cu; measurement build only) ---------------------------------------------------
# binned: chunks of 256 consecutive queries of one head; the grad_value contributions of the COARSE levels (decided on
# the device: coarsest first while their cells fit 704 and their samples fit 8 per query) are combined in shared memory.
# tile: a persistent grid over work items (image, head, tile).  When the queries ARE the pixel pyramid (Lq == sum H*W:
# the encoder) a tile is a 12 x 16 block of the largest level plus the coarser levels' pixels of the same image region,
# and the backward combines the contributions of EVERY level inside per-level windows (tile + 6 pixel halo); other
# query sets run in chunks of consecutive queries with whole-level windows.  bwd_variant 21 / 20 force the binned / tile
# backward for any Lq, fwd_variant 12 the tile forward.
def _encoder_case(seed, shapes, N, M, D, P, sigma_px=2.0):
    """queries = the pixel pyramid; locations = own pixel centre on every level + N(0, sigma_px) offsets
    (SURVEY.md 8d 'model-like'): what MonoDETR's encoder produces (reference depthaware_transformer.py:363-376)"""
    from monosowa_b200.workloads import encoder_reference_points
    g = torch.Generator().manual_seed(seed)
    sh, lsi = _levels(shapes)
    S, L = int(sh.prod(1).sum()), len(shapes)
    value = torch.randn(N, S, M, D, generator=g, dtype=torch.float64)
    ref = encoder_reference_points(shapes, "cpu", torch.float64)[None].expand(N, -1, -1, -1)
    wh = torch.stack([sh[:, 1], sh[:, 0]], -1).double()
    off = torch.randn(N, S, M, L, P, 2, generator=g, dtype=torch.float64) * sigma_px
    loc = (ref[:, :, None, :, None, :] + off / wh[None, None, None, :, None, :]).contiguous()
    attn = torch.softmax(torch.randn(N, S, M, L * P, generator=g, dtype=torch.float64), -1).view(N, S, M, L, P)
    grad_out = torch.randn(N, S, M * D, generator=g, dtype=torch.float64)
    return value, sh, lsi, loc, attn, grad_out


GRID_CASES = [
    # (shapes, N, M, D, P, sigma_px)                                  queries = the pyramid (grid mode)
    ([(24, 80), (12, 40), (6, 20), (3, 10)], 2, 8, 32, 4, 2.0),      # MonoDETR layout at half size: 2 x 5 tiles
    ([(24, 80), (12, 40), (6, 20), (3, 10)], 1, 2, 32, 4, 9.0),      # large offsets: many samples leave the windows (strays)
    ([(13, 37), (7, 19), (4, 10)], 2, 3, 32, 4, 2.0),                # not dyadic, ragged tiles, M = 3
    ([(30, 30)], 1, 2, 32, 4, 1.5),                                  # a single level
    ([(3, 10), (24, 80), (6, 20), (12, 40)], 1, 2, 32, 4, 2.0),      # levels not ordered by size: tiling base = level 1
    ([(6, 4), (3, 2)], 1, 2, 32, 2, 1.0),                            # ops/test.py pyramid: one tile
    ([(24, 40), (12, 20), (6, 10)], 2, 4, 16, 4, 2.0),               # D = 16: 4-lane groups, three batches of 4 samples
    ([(24, 40), (12, 20)], 1, 2, 64, 4, 2.0),                        # D = 64: 16-lane groups, rounds of 128 queries
    ([(24, 40), (12, 20), (6, 10)], 1, 4, 32, 3, 2.0),               # P = 3: batches straddle levels, ragged last batch
    ([(20, 33), (10, 17)], 1, 2, 32, 16, 2.0),                       # P = 16: a level spans two batches
    ([(40, 48), (40, 48), (40, 48)], 1, 1, 32, 4, 2.0),              # equal levels: 576 queries per tile -> three rounds
]

LINEAR_CASES = [
    # (shapes, N, M, D, Lq, P)                                        queries are NOT the pyramid (linear mode)
    ([(12, 40), (6, 20), (3, 10), (2, 5)], 2, 8, 32, 300, 4),        # coarse levels windowed whole; two chunks, ragged
    ([(9, 7), (5, 4), (3, 3)], 2, 4, 32, 519, 3),                    # P = 3, three chunks
    ([(9, 7), (5, 4)], 2, 4, 16, 700, 8),                            # D = 16, P = 8
    ([(9, 7), (5, 4)], 2, 4, 64, 150, 4),                            # D = 64 (128-query chunks)
    ([(16, 16)], 1, 1, 32, 257, 4),                                  # single level, one query in the last chunk
    ([(40, 40), (30, 30)], 1, 2, 32, 300, 4),                        # no level fits the cell budget: direct reductions only
    ([(9, 7), (1, 1)], 2, 3, 32, 260, 4),                            # 1x1 coarsest level: four cells, one hot row
    ([(6, 4), (3, 2)], 1, 2, 32, 2, 2),                              # two queries
]


def _tile_vs_oracle_and_record(msda, case_inputs, dtype, label, variants=(12, 20)):
    """binned / tile kernels against the fp64 oracle, and against the record kernels: out, grad_loc and grad_attn come
    from identical per-query arithmetic (bitwise equal), grad_value differs only in summation order."""
    value, sh, lsi, loc, attn, grad_out = case_inputs
    L = msda._lib
    ct = torch.float32
    try:
        L.set_tuning("fwd_variant", variants[0]); L.set_tuning("bwd_variant", variants[1])
        check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, dtype, label=label)
        b = run_ours(msda, value.to(dtype), sh, lsi, loc.to(ct), attn.to(ct), grad_out.to(dtype))
        # the record backward with the same lane layout: the same partial sums in the same order (the binned kernel takes
        # 8 channels per lane at D = 64 like the record kernel, the tile kernels always 4)
        L.set_tuning("fwd_variant", 11); L.set_tuning("bwd_variant", 11)
        if variants[1] == 20:
            L.set_tuning("bwd_pipe", 4)
        a = run_ours(msda, value.to(dtype), sh, lsi, loc.to(ct), attn.to(ct), grad_out.to(dtype))
    finally:
        L.set_tuning("fwd_variant", -1); L.set_tuning("bwd_variant", -1); L.set_tuning("bwd_pipe", -1)
    assert torch.equal(a[0], b[0]), "forward differs from the record kernel"
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]), "grad_loc / grad_attn differ from the record kernel"
    assert O.rel_l2(b[1], a[1]) < (1e-5 if dtype == torch.float32 else 1e-2)


KERNEL_FAMILIES = [pytest.param((11, 21), id="binned"), pytest.param((12, 20), id="tile")]


@pytest.mark.parametrize("variants", KERNEL_FAMILIES)
@pytest.mark.parametrize("case", GRID_CASES, ids=lambda c: f"M{c[2]}D{c[3]}P{c[4]}L{len(c[0])}s{c[5]}")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_long_query_kernels_on_pixel_pyramids(msda, case, dtype, variants):
    if variants[1] == 20:
        _needs_ab(msda)
    shapes, N, M, D, P, sigma = case
    _tile_vs_oracle_and_record(msda, _encoder_case(700 + D + P, shapes, N, M, D, P, sigma), dtype, f"grid {case}", variants)


@pytest.mark.parametrize("variants", KERNEL_FAMILIES)
@pytest.mark.parametrize("case", LINEAR_CASES, ids=lambda c: f"M{c[2]}D{c[3]}Lq{c[4]}P{c[5]}L{len(c[0])}")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_long_query_kernels_on_arbitrary_query_sets(msda, case, dtype, variants):
    if variants[1] == 20:
        _needs_ab(msda)
    shapes, N, M, D, Lq, P = case
    inputs = _random_case(300 + D + Lq, shapes, N, M, D, Lq, P, spread=1.4, shift=-0.2)
    _tile_vs_oracle_and_record(msda, inputs, dtype, f"linear {case}", variants)


@pytest.mark.parametrize("variants", KERNEL_FAMILIES)
def test_long_query_kernels_uniform_locations_on_a_pyramid(msda, variants):
    """locations that ignore the query's own position: for the tile kernel nearly every fine-level sample is a stray"""
    if variants[1] == 20:
        _needs_ab(msda)
    shapes = [(24, 80), (12, 40), (6, 20), (3, 10)]
    S = sum(h * w for h, w in shapes)
    inputs = _random_case(811, shapes, 1, 4, 32, S, 4, spread=1.3, shift=-0.15)
    _tile_vs_oracle_and_record(msda, inputs, torch.float32, "uniform on pyramid", variants)


def test_binned_backward_is_the_default_for_long_query_sets(msda, cuda_device):
    """Lq >= 1024 with enough (image, head, chunk) work items takes the binned kernel without any tuning; hot coarse
    pixels (every query samples the same corner of the coarsest levels) stress the per-cell counters: 2048 entries in
    one cell."""
    assert msda._lib.describe("forward", torch.float32, 16, 8, 32, 4, 4, 10200) == "fwd_rec_f32"
    assert msda._lib.describe("backward", torch.float32, 16, 8, 32, 4, 4, 10200) == "bwd_bin_f32"
    assert msda._lib.describe("backward", torch.float32, 16, 8, 32, 4, 4, 550) == "bwd_rec_f32"
    shapes, N, M, D, Lq, P = [(12, 40), (6, 20), (3, 10), (2, 5)], 16, 40, 32, 1100, 4
    assert msda._lib.describe("backward", torch.float32, N, M, D, len(shapes), P, Lq) == "bwd_bin_f32"
    value, sh, lsi, loc, attn, grad_out = _random_case(41, shapes, N, M, D, Lq, P, spread=1.2, shift=-0.1)
    loc[:, :, :, 2:, :, :] = 0.26                                 # all coarse samples in one cell
    check_against_oracle(msda, value, sh, lsi, loc, attn, grad_out, torch.float32, label="hot cell")
    L = msda._lib
    a = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    try:
        L.set_tuning("fwd_variant", 11); L.set_tuning("bwd_variant", 11)
        b = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    finally:
        L.set_tuning("fwd_variant", -1); L.set_tuning("bwd_variant", -1)
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and O.rel_l2(a[1], b[1]) < 1e-5


def test_nan_behind_outside_samples_does_not_leak(msda, cuda_device):
    """The reference never touches `value` for a sample outside the (-1,H)x(-1,W) window (cuh:288).  Neither do
    these kernels: a NaN in pixel 0 of every head and level must not reach queries whose samples there are outside."""
    shapes, N, M, D, Lq, P = [(6, 8), (3, 4)], 1, 2, 32, 40, 4
    value, sh, lsi, loc, attn, grad_out = _random_case(97, shapes, N, M, D, Lq, P, spread=0.5, shift=0.3)
    loc[:, ::2, :, :, 0] = 1.7                                     # every second query: one sample per level far outside
    for dtype in (torch.float32, torch.bfloat16):
        v = value.to(dtype).clone()
        clean = run_ours(msda, v, sh, lsi, loc.float(), attn.float(), grad_out.to(dtype))
        v[:, 0] = float("nan"); v[:, int(lsi[1])] = float("nan")   # pixel (0,0) of both levels
        for fv, bv in ((11, 11), (11, 21)) + (((12, 20),) if msda._lib.has_ab_flavours() else ()):
            msda._lib.set_tuning("fwd_variant", fv); msda._lib.set_tuning("bwd_variant", bv)
            try:
                got = run_ours(msda, v, sh, lsi, loc.float(), attn.float(), grad_out.to(dtype))
            finally:
                msda._lib.set_tuning("fwd_variant", -1); msda._lib.set_tuning("bwd_variant", -1)
            # locations in [0.3, 0.8] never touch pixel (0,0) with a non-zero weight... unless a corner is pixel 0:
            touches = torch.isnan(got[0]).reshape(N, Lq, -1).any(-1)
            ref_touch = torch.isnan(O.forward_c(v.double(), sh, lsi, loc, attn)).reshape(N, Lq, -1).any(-1)
            assert torch.equal(touches, ref_touch), f"NaN pattern differs from the oracle ({dtype}, variants {fv}/{bv})"
            ok = ~ref_touch
            assert torch.equal(got[0][ok], clean[0][ok])
            assert torch.isfinite(got[2][ok]).all() and torch.isfinite(got[3][ok]).all()


def test_edge_locations_exact_borders(msda):
    g = load_golden("op", "d32_edges")
    for dtype in (torch.float64, torch.float32):
        check_against_oracle(msda, g["value"], g["shapes"], g["level_start_index"], g["loc"], g["attn"],
                             g["grad_out"], dtype, label="edges")


def test_all_samples_out_of_range(msda, cuda_device):
    value, sh, lsi, loc, attn, grad_out = _random_case(9, [(5, 6), (3, 3)], 2, 8, 32, 21, 4)
    loc = loc.abs() + 1.3                                            # everything beyond the right/bottom edge
    out, gv, gl, ga = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    assert (out == 0).all() and (gv == 0).all() and (gl == 0).all() and (ga == 0).all()


def test_empty_and_degenerate_shapes(msda, cuda_device):
    sh, lsi = _levels([(4, 4), (2, 2)])
    dev = cuda_device
    value = torch.randn(2, 20, 8, 32, device=dev)
    loc = torch.rand(2, 0, 8, 2, 4, 2, device=dev)
    attn = torch.rand(2, 0, 8, 2, 4, device=dev)
    v = value.clone().requires_grad_(True)
    out = msda.MSDeformAttnFunction.apply(v, sh.to(dev), lsi.to(dev), loc, attn, 64)
    assert out.shape == (2, 0, 256)
    out.sum().backward()
    assert (v.grad == 0).all()
    # batch 0
    out0 = msda.MSDeformAttnFunction.apply(value[:0], sh.to(dev), lsi.to(dev), torch.rand(0, 3, 8, 2, 4, 2, device=dev),
                                           torch.rand(0, 3, 8, 2, 4, device=dev), 64)
    assert out0.shape == (0, 3, 256)


# --------------------------------------------------------------------------------------------
# 4. full-size workloads (BASELINE.json configs)
# --------------------------------------------------------------------------------------------
def test_config0_shape_vs_oracle_f32(msda):
    """configs[0] shape (batch 2, 10200 queries): full comparison, the oracle takes seconds."""
    from monosowa_b200 import workloads as W
    wl = W.config(0)
    d = W.make_inputs(wl)
    e = check_against_oracle(msda, d["value"].double(), d["shapes"], d["lsi"], d["loc"].double(),
                             d["attn"].double(), d["grad_out"].double(), torch.float32, label=wl.name)
    print("config0 errors", e)


@pytest.mark.parametrize("mode", ["model", "uniform"])
def test_config1_full_size_f32(msda, cuda_device, mode):
    """configs[1] (the headline: batch 16, fp32).  Oracle comparison on one image of the batch plus
    size-independent properties on the whole batch: linearity in value, the adjoint identity
    <f(v), g> = <v, grad_value(g)>, <f(v; a'), g> = <a', grad_attn(g)>, bitwise-repeatable forward."""
    from monosowa_b200 import workloads as W
    wl = W.config(1, loc_mode=mode)
    d = W.make_inputs(wl, device=cuda_device)
    F = msda.MSDeformAttnFunction.apply
    v = d["value"].requires_grad_(True); l = d["loc"].requires_grad_(True); a = d["attn"].requires_grad_(True)
    out = F(v, d["shapes"], d["lsi"], l, a, 64)
    out.backward(d["grad_out"])
    # (a) one image against the fp64 oracle
    i = 5
    sl = lambda t: t[i:i + 1].detach().cpu().double()
    ref_out = O.forward_c(sl(v), d["shapes"].cpu(), d["lsi"].cpu(), sl(l), sl(a))
    rgv, rgl, rga = O.backward_c(sl(v), d["shapes"].cpu(), d["lsi"].cpu(), sl(l), sl(a), sl(d["grad_out"]))
    assert O.rel_l2(out[i:i + 1], ref_out) < 1e-5
    assert O.max_abs_over_max(out[i:i + 1], ref_out) < 2e-5
    assert O.rel_l2(v.grad[i:i + 1], rgv) < 1e-4
    assert O.rel_l2(a.grad[i:i + 1], rga) < 1e-4
    keep = ~O.pixel_boundary_mask(sl(l), d["shapes"].cpu(), eps_px=1e-4)
    assert O.rel_l2(l.grad[i:i + 1].cpu()[keep], rgl[keep]) < 1e-4
    # (b) properties over the whole batch (fp64 accumulation of the inner products)
    with torch.no_grad():
        v2 = torch.randn_like(v)
        o2 = F(v2, d["shapes"], d["lsi"], l, a, 64)
        lin = F(2.5 * v - v2, d["shapes"], d["lsi"], l, a, 64)
        assert O.rel_l2(lin, 2.5 * out - o2) < 1e-5
        lhs = (o2.double() * d["grad_out"].double()).sum().item()
        rhs = (v2.double() * v.grad.double()).sum().item()
        scale = (o2.double().norm() * d["grad_out"].double().norm()).item()
        assert abs(lhs - rhs) < 1e-5 * scale
        a2 = torch.rand_like(a)
        o3 = F(v, d["shapes"], d["lsi"], l, a2, 64)
        lhs = (o3.double() * d["grad_out"].double()).sum().item()
        rhs = (a2.double() * a.grad.double()).sum().item()
        scale = (o3.double().norm() * d["grad_out"].double().norm()).item()
        assert abs(lhs - rhs) < 1e-5 * scale
        assert torch.equal(F(v, d["shapes"], d["lsi"], l, a, 64), out)


@pytest.mark.parametrize("lq", [50, 550])
def test_config2_decoder_bf16(msda, lq):
    """configs[2]: decoder cross-attention, 50 (eval) / 550 (train) queries, bf16 value."""
    from monosowa_b200 import workloads as W
    wl = W.config(2, num_queries=lq)                       # batch 16, as BASELINE.json states
    d = W.make_inputs(wl)
    # short query sets scatter straight into the bf16 gradient (REDG.E.ADD.BF16x4, no fp32 buffer, no narrowing pass)
    assert msda._lib.lib.msda_backward_bf16_scratch_bytes(wl.batch, wl.S, wl.heads, wl.head_dim, wl.L, lq, wl.points, 1) == 0
    e = check_against_oracle(msda, d["value"].double(), d["shapes"], d["lsi"], d["loc"].double(), d["attn"].double(),
                             d["grad_out"].double(), torch.bfloat16, label=wl.name)
    print(f"config2 Lq={lq} bf16 errors {e}")
    assert e["gv"] < 8e-3, "per-addition bf16 rounding of grad_value must stay inside the 1e-2 contract (measured 5.3e-3 at Lq=550)"


def test_config4_large_image_shapes(msda):
    """configs[4] shapes (KITTI-360, Waymo 1280x1920, 640x960), batch 4 as SURVEY.md 8d states, fp32 + bf16
    (the C oracle needs about a minute for the Waymo shape)."""
    from monosowa_b200 import workloads as W
    for wl in W.sweep_config5(batch=4):
        d = W.make_inputs(wl)
        check_against_oracle(msda, d["value"].double(), d["shapes"], d["lsi"], d["loc"].double(),
                             d["attn"].double(), d["grad_out"].double(), wl.dtype, label=wl.name)


def test_beyond_int32_indexing(msda, cuda_device):
    """Maximum sizes: value with 2^31 + 4096 elements (8.6 GB).  The reference kernels index with int32
    (ms_deform_im2col_cuda.cuh:47-53) and would overflow here; this library switches to its 64-bit generic
    kernels.  All samples of level 0 fall into its bottom-right corner, so the expected result is the oracle
    on a cropped problem (last 8 rows x 12 columns of level 0 + the whole level 1) with re-normalised
    locations; everything outside the crop must receive exactly zero gradient."""
    dev = cuda_device
    H, W, M, D, Lq, P, CH, CW = 2048, 4096, 8, 32, 24, 3, 8, 12
    g = torch.Generator(device=dev).manual_seed(4242)
    sh = torch.tensor([[H, W], [4, 4]], device=dev); lsi = torch.tensor([0, H * W], device=dev)
    S = H * W + 16
    assert S * M * D >= 2 ** 31
    value = torch.randn(1, S, M, D, generator=g, device=dev)
    px = torch.rand(1, Lq, M, 1, P, generator=g, device=dev) * 8.8 + (W - 9)        # up to 0.3 px beyond the right edge
    py = torch.rand(1, Lq, M, 1, P, generator=g, device=dev) * 4.8 + (H - 5)
    loc0 = torch.stack([(px + 0.5) / W, (py + 0.5) / H], -1)
    loc1 = torch.rand(1, Lq, M, 1, P, 2, generator=g, device=dev) * 1.4 - 0.2
    loc = torch.cat([loc0, loc1], 3).contiguous().requires_grad_(True)
    attn = torch.softmax(torch.randn(1, Lq, M, 2 * P, generator=g, device=dev), -1).view(1, Lq, M, 2, P).requires_grad_(True)
    grad_out = torch.randn(1, Lq, M * D, generator=g, device=dev)
    v = value.requires_grad_(True)
    assert msda._lib.describe("forward", torch.float32, 1, M, D, 2, P, Lq) == "fwd_rec_f32"   # the shape alone would vectorise
    out = msda.MSDeformAttnFunction.apply(v, sh, lsi, loc, attn, 64)
    out.backward(grad_out)
    torch.cuda.synchronize()
    # cropped fp64 problem, pixel coordinates taken exactly as the fp32 kernel forms them
    l32 = loc.detach()
    pxk = (l32[..., 0, :, 0] * W).float() - 0.5
    pyk = (l32[..., 0, :, 1] * H).float() - 0.5
    locc0 = torch.stack([(pxk.double() - (W - CW) + 0.5) / CW, (pyk.double() - (H - CH) + 0.5) / CH], -1)[:, :, :, None]
    locc = torch.cat([locc0, l32[..., 1:2, :, :].double()], 3).cpu()
    lvl0 = value.detach()[0, :H * W].view(H, W, M, D)[H - CH:, W - CW:].reshape(1, CH * CW, M, D)
    vcrop = torch.cat([lvl0, value.detach()[:, H * W:]], 1).double().cpu()
    shc, lsic = torch.tensor([[CH, CW], [4, 4]]), torch.tensor([0, CH * CW])
    ref_out = O.forward_c(vcrop, shc, lsic, locc, attn.detach().double().cpu())
    rgv, rgl, rga = O.backward_c(vcrop, shc, lsic, locc, attn.detach().double().cpu(), grad_out.double().cpu())
    assert O.rel_l2(out, ref_out) < 1e-5
    assert O.rel_l2(attn.grad, rga) < 1e-4
    gv = v.grad[0]
    gcrop = torch.cat([gv[:H * W].view(H, W, M, D)[H - CH:, W - CW:].reshape(CH * CW, M, D), gv[H * W:]], 0)
    assert O.rel_l2(gcrop, rgv[0]) < 1e-4
    assert abs(gv.double().abs().sum().item() - gcrop.double().abs().sum().item()) < 1e-6 * gcrop.double().abs().sum().item()
    # d/dloc of level 0 differs from the cropped problem by the normalisation ratio (W/CW, H/CH)
    keep = ~O.pixel_boundary_mask(locc, shc, eps_px=2e-3)            # fp32 pixel coordinates near 4096 resolve 2.4e-4 px
    scale = torch.tensor([W / CW, H / CH], dtype=torch.float64)
    rgl_full = rgl.clone(); rgl_full[..., 0, :, :] *= scale
    assert O.rel_l2(loc.grad.cpu()[keep], rgl_full[keep]) < 1e-4


# --------------------------------------------------------------------------------------------
# 5. against the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref)
# --------------------------------------------------------------------------------------------
@pytest.mark.skipif(not O.ref_cuda_available(), reason="oracle/_ref not built (needs /root/reference at build time)")
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("loc_mode", ["model", "init"])
def test_matches_reference_cuda_kernels(msda, cuda_device, dtype, loc_mode):
    """`init`: the module's initial offsets (whole pixels from a pixel centre, ms_deform_attn.py:106-115) put EVERY
    sample on a pixel boundary, where d out / d loc jumps: only the reference's exact coordinate arithmetic -- one
    fused multiply-add, see make_tap in msda_common.cuh -- reproduces its floor() decisions."""
    from monosowa_b200 import workloads as W
    wl = W.config(0, batch=2, dtype=dtype, loc_mode=loc_mode)
    d = W.make_inputs(wl, device=cuda_device)
    ours = torch.ops.msda.forward(d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"], 64)
    ref = O.ref_cuda_forward(d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
    gv, gl, ga = torch.ops.msda.backward(d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"], d["grad_out"], 64)
    rgv, rgl, rga = O.ref_cuda_backward(d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"], d["grad_out"])
    torch.cuda.synchronize()
    t = 1e-6 if dtype == torch.float32 else 1e-14
    # identical fp32 floor() decisions by construction -> no boundary masking needed here
    assert O.rel_l2(ours, ref) < t
    assert O.rel_l2(gv, rgv) < 10 * t and O.rel_l2(gl, rgl) < 10 * t and O.rel_l2(ga, rga) < 10 * t


# --------------------------------------------------------------------------------------------
# 6. API / error behaviour / streams / graphs
# --------------------------------------------------------------------------------------------
def test_error_behaviour_matches_reference(msda, cuda_device):
    value, sh, lsi, loc, attn, _ = _random_case(10, [(5, 6), (3, 3)], 4, 8, 32, 9, 4)
    dev = cuda_device
    v, l, a = value.float().to(dev), loc.float().to(dev), attn.float().to(dev)
    sh, lsi = sh.to(dev), lsi.to(dev)
    F = msda.MSDeformAttnFunction.apply
    with pytest.raises(RuntimeError, match="contiguous"):                        # ms_deform_attn_cuda.cu:28
        F(v.transpose(1, 2).contiguous().transpose(1, 2), sh, lsi, l, a, 64)
    with pytest.raises(RuntimeError, match="im2col_step"):                       # ms_deform_attn_cuda.cu:52
        F(v, sh, lsi, l, a, 3)
    with pytest.raises(NotImplementedError):                                     # ms_deform_attn.h:38
        F(v.cpu(), sh.cpu(), lsi.cpu(), l.cpu(), a.cpu(), 64)
    with pytest.raises(RuntimeError):
        F(v.half(), sh, lsi, l, a, 64)
    with pytest.raises(RuntimeError, match="inconsistent"):
        F(v, sh, lsi, l[:, :, :4].contiguous(), a, 64)
    assert F(v, sh, lsi, l, a, 2).shape == (4, 9, 256)                           # 4 % 2 == 0 is fine


def test_c_abi_direct_call_and_error_codes(msda, cuda_device):
    """Call the exported symbols with raw pointers, as a non-Python host would."""
    lib = msda._lib.lib
    value, sh, lsi, loc, attn, _ = _random_case(11, [(5, 6), (3, 3)], 1, 8, 32, 9, 4)
    dev = cuda_device
    v, l, a = value.float().to(dev), loc.float().to(dev), attn.float().to(dev)
    shd, lsid = sh.to(dev), lsi.to(dev)
    out = torch.full((1, 9, 256), float("nan"), device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n0 = msda._lib.launch_count()
    rc = lib.msda_forward_f32(p(v), p(shd), p(lsid), p(l), p(a), p(out), 1, 39, 8, 32, 2, 9, 4, st)
    torch.cuda.synchronize()
    assert rc == 0 and msda._lib.last_error() == "" and msda._lib.launch_count() == n0 + 1
    ref = O.forward_c(v.cpu().double(), sh, lsi, l.cpu().double(), a.cpu().double())
    assert O.rel_l2(out, ref) < 1e-5
    assert lib.msda_forward_f32(None, p(shd), p(lsid), p(l), p(a), p(out), 1, 39, 8, 32, 2, 9, 4, st) == -1
    assert "NULL" in msda._lib.last_error()
    assert lib.msda_forward_f32(p(v), p(shd), p(lsid), p(l), p(a), p(out), 1, 39, 8, 32, 99, 9, 4, st) == -2
    assert lib.msda_forward_f32(p(v), p(shd), p(lsid), p(l), p(a), p(out), -1, 39, 8, 32, 2, 9, 4, st) == -2
    # a 4-byte-aligned but not 16-byte-aligned view must still work (generic kernels)
    buf = torch.zeros(v.numel() + 1, device=dev)
    buf[1:].copy_(v.flatten())
    out2 = torch.empty_like(out)
    rc = lib.msda_forward_f32(ctypes.c_void_p(buf.data_ptr() + 4), p(shd), p(lsid), p(l), p(a), p(out2),
                              1, 39, 8, 32, 2, 9, 4, st)
    torch.cuda.synchronize()
    assert rc == 0 and O.rel_l2(out2, out) < 2e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float64])
@pytest.mark.parametrize("shape", [([(5, 7), (3, 4), (1, 1)], 2, 8, 32, 37, 4), ([(4, 3)], 1, 3, 16, 9, 2), ([(6, 5), (2, 2)], 1, 2, 24, 5, 3)])
@pytest.mark.parametrize("bwd_variant", [-1, 21, 20])
def test_guard_bands_no_out_of_bounds_access(msda, cuda_device, dtype, shape, bwd_variant):
    if bwd_variant == 20:
        _needs_ab(msda)
    msda._lib.set_tuning("bwd_variant", bwd_variant)
    try:
        _guard_bands(msda, cuda_device, dtype, shape)
    finally:
        msda._lib.set_tuning("bwd_variant", -1)


def _guard_bands(msda, cuda_device, dtype, shape):
    """compute-sanitizer is closed on this GPU pool, so bounds are checked the hard way: every tensor
    is carved out of a larger buffer whose surroundings are NaN (inputs) or a sentinel (outputs), the
    locations hammer the borders, and the kernels are called through the raw C ABI.  An out-of-range
    read poisons the results with NaN (even at weight 0); an out-of-range write breaks a sentinel."""
    shapes, N, M, D, Lq, P = shape
    value, sh, lsi, loc, attn, grad_out = _random_case(31, shapes, N, M, D, Lq, P, spread=1.6, shift=-0.3)
    border = torch.tensor([-1e-4, 0.0, 1e-4, 0.5, 1.0 - 1e-4, 1.0, 1.0 + 1e-4, -0.2, 1.2], dtype=torch.float64)
    pick = torch.randint(0, len(border), loc.shape, generator=torch.Generator().manual_seed(5))
    loc = torch.where(torch.rand(loc.shape, generator=torch.Generator().manual_seed(6)) < 0.5, border[pick], loc)
    dev, ct = cuda_device, (torch.float64 if dtype == torch.float64 else torch.float32)
    gvt = dtype                                                    # grad_value has value's dtype
    GUARD = 4096                                                   # elements on each side

    def carve(t, fill):
        buf = torch.full((t.numel() + 2 * GUARD,), fill, dtype=t.dtype, device=dev)
        view = buf[GUARD:GUARD + t.numel()].view(t.shape)
        return buf, view

    bufs = {}
    ins = dict(value=value.to(dtype), loc=loc.to(ct), attn=attn.to(ct), grad_out=grad_out.to(dtype))
    views = {}
    for k, t in ins.items():
        bufs[k], views[k] = carve(t.to(dev), float("nan"))
        views[k].copy_(t)
    outs = dict(out=torch.empty(N, Lq, M * D, dtype=dtype), gv=torch.empty(value.shape, dtype=gvt),
                gl=torch.empty(loc.shape, dtype=ct), ga=torch.empty(attn.shape, dtype=ct))
    SENT = 4096.0                                                  # exactly representable in bf16
    for k, t in outs.items():
        bufs[k], views[k] = carve(t.to(dev), SENT)
    lib = msda._lib.lib
    sfx = {torch.float32: "f32", torch.bfloat16: "bf16", torch.float64: "f64"}[dtype]
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    shd, lsid = sh.to(dev), lsi.to(dev)
    S, L = int(sh.prod(1).sum()), len(shapes)
    rc = getattr(lib, "msda_forward_" + sfx)(p(views["value"]), p(shd), p(lsid), p(views["loc"]), p(views["attn"]),
                                             p(views["out"]), N, S, M, D, L, Lq, P, st)
    assert rc == 0, msda._lib.last_error()
    extra = []
    if dtype == torch.bfloat16:                                    # fp32 accumulation buffer, when this shape needs one
        nb = lib.msda_backward_bf16_scratch_bytes(N, S, M, D, L, Lq, P, 0)     # 0: the carved views are only 4-byte aligned
        outs["scratch"] = torch.empty(max(nb // 4, 1), dtype=torch.float32)
        bufs["scratch"], views["scratch"] = carve(outs["scratch"].to(dev), SENT)
        extra = [p(views["scratch"])]
    rc = getattr(lib, "msda_backward_" + sfx)(p(views["value"]), p(shd), p(lsid), p(views["loc"]), p(views["attn"]),
                                              p(views["grad_out"]), p(views["gv"]), p(views["gl"]), p(views["ga"]), *extra,
                                              N, S, M, D, L, Lq, P, st)
    assert rc == 0, msda._lib.last_error()
    torch.cuda.synchronize()
    for k in outs:
        assert torch.isfinite(views[k].float()).all(), f"{k}: NaN leaked in -> out-of-range read"
        b = bufs[k].float()
        assert (b[:GUARD] == SENT).all() and (b[-GUARD:] == SENT).all(), f"{k}: guard band overwritten"
    args = (ins["value"].double(), sh, lsi, ins["loc"].double(), ins["attn"].double())
    tol = TOL[dtype]
    assert O.rel_l2(views["out"], O.forward_c(*args)) <= tol["fwd"]
    rgv, rgl, rga = O.backward_c(*args, ins["grad_out"].double())
    assert O.rel_l2(views["gv"], rgv) <= tol["gv"]
    assert O.rel_l2(views["ga"], rga) <= tol["ga"]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("per_chunk", [1, 2, 3, 5])
def test_host_buffer_step_matches_device_path(msda, cuda_device, dtype, per_chunk):
    """msda_host_step_*: pinned host tensors in, pinned host results out, pipelined over the batch (7 images:
    more chunks than pipeline stages, a ragged last chunk).  Same kernels as the device path: out, grad_loc and
    grad_attn bitwise equal, grad_value up to summation order."""
    shapes, N, M, D, Lq, P = [(12, 40), (6, 20), (3, 10), (2, 5)], 7, 8, 32, 131, 4
    value, sh, lsi, loc, attn, grad_out = _random_case(55, shapes, N, M, D, Lq, P)
    v, g = value.to(dtype), grad_out.to(dtype)
    l, a = loc.float(), attn.float()
    want = run_ours(msda, v, sh, lsi, l, a, g)
    dev = cuda_device
    out, gv, gl, ga = msda.host_step(v.pin_memory(), sh.to(dev), lsi.to(dev), l.pin_memory(), a.pin_memory(),
                                     g.reshape(N, Lq, M * D).pin_memory(), images_per_chunk=per_chunk)
    assert not out.is_cuda and out.dtype == dtype and gv.dtype == dtype
    assert torch.equal(out, want[0]) and torch.equal(gl, want[2]) and torch.equal(ga, want[3])
    assert O.rel_l2(gv, want[1]) < (1e-5 if dtype == torch.float32 else 1e-2)
    # a second call reuses the cached workspace and the caller's result buffers; a deeper ring (5 stages) changes nothing
    res2 = msda.host_step(v.pin_memory(), sh.to(dev), lsi.to(dev), l.pin_memory(), a.pin_memory(),
                          g.reshape(N, Lq, M * D).pin_memory(), images_per_chunk=per_chunk, results=(out, gv, gl, ga), stages=5)
    assert res2[0] is out and torch.equal(out, want[0]) and torch.equal(gl, want[2])
    with pytest.raises(RuntimeError):
        msda.host_step(v.to(dev), sh.to(dev), lsi.to(dev), l, a, g.reshape(N, Lq, M * D))       # device tensor: wrong entry point
    with pytest.raises(NotImplementedError):
        msda.host_step(v, sh, lsi, l, a, g.reshape(N, Lq, M * D))                                # no device anywhere


def test_side_stream_and_cuda_graph(msda, cuda_device):
    from monosowa_b200 import workloads as W
    d = W.make_inputs(W.config(0, batch=1), device=cuda_device)
    args = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
    want = torch.ops.msda.forward(*args, 64)
    wgv, wgl, wga = torch.ops.msda.backward(*args, d["grad_out"], 64)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        got = torch.ops.msda.forward(*args, 64)
    s.synchronize()
    assert torch.equal(got, want)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):                        # no host sync / host metadata reads inside
        g_out = torch.ops.msda.forward(*args, 64)
        g_gv, g_gl, g_ga = torch.ops.msda.backward(*args, d["grad_out"], 64)
    g_out.zero_(); g_gv.fill_(7.0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(g_out, want)
    assert O.rel_l2(g_gv, wgv) < 1e-5 and torch.equal(g_gl, wgl) and torch.equal(g_ga, wga)


def test_registered_op_autograd_and_fake(msda, cuda_device):
    value, sh, lsi, loc, attn, grad_out = _random_case(12, [(5, 6), (3, 3)], 2, 8, 32, 9, 4)
    dev = cuda_device
    v = value.float().to(dev).requires_grad_(True)
    l = loc.float().to(dev).requires_grad_(True)
    a = attn.float().to(dev).requires_grad_(True)
    out = torch.ops.msda.forward(v, sh.to(dev), lsi.to(dev), l, a, 64)       # autograd registered on the op itself
    out.backward(grad_out.float().to(dev))
    ref = run_ours(msda, value.float(), sh, lsi, loc.float(), attn.float(), grad_out.float())
    assert torch.equal(out.detach().cpu(), ref[0])
    assert O.rel_l2(v.grad, ref[1]) < 1e-5 and torch.equal(l.grad.cpu(), ref[2]) and torch.equal(a.grad.cpu(), ref[3])
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        fv = torch.empty(2, 39, 8, 32, device="cuda")
        fo = torch.ops.msda.forward(fv, torch.empty(2, 2, dtype=torch.long, device="cuda"),
                                    torch.empty(2, dtype=torch.long, device="cuda"),
                                    torch.empty(2, 9, 8, 2, 4, 2, device="cuda"), torch.empty(2, 9, 8, 2, 4, device="cuda"), 64)
        assert fo.shape == (2, 9, 256)


# --------------------------------------------------------------------------------------------
# 6b. fused pre-processing: softmax + ref + offset/(W,H) inside the kernels
# --------------------------------------------------------------------------------------------
FUSED_CASES = [
    ([(12, 40), (6, 20), (3, 10), (2, 5)], 2, 8, 32, 131, 4),
    ([(9, 7), (5, 4), (3, 3)], 2, 4, 32, 19, 3),              # L*P = 9: ragged second batch
    ([(9, 7), (5, 4)], 3, 3, 16, 17, 2),                      # 4-lane groups, one batch
    ([(9, 7), (5, 4), (3, 3), (2, 2)], 1, 2, 64, 9, 4),       # 16-lane groups
    ([(9, 7), (5, 4), (3, 3), (2, 2)], 1, 2, 16, 9, 4),       # L*P = 16 = 4 batches of 4 lanes
    ([(9, 7), (5, 4), (3, 3)], 1, 2, 32, 9, 8),               # L*P = 24 > 4 batches of the 4-lane (8-channel) groups: 8-lane flavour
    ([(9, 7), (5, 4), (3, 3), (2, 2), (1, 1)], 1, 2, 64, 9, 8),   # L*P = 40: D = 64 falls back to 16-lane groups
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=lambda c: f"M{c[2]}D{c[3]}Lq{c[4]}P{c[5]}L{len(c[0])}")
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("family", [(11, 11), (11, 21), (12, 20)], ids=["record", "binned", "tile"])
@pytest.mark.parametrize("ref_dim", [2, 6])
def test_fused_preprocessing_matches_unfused_and_oracle(msda, cuda_device, case, dtype, family, ref_dim):
    """the fused flavours of the record kernels, the binned backward and (measurement build) the tile kernels, with
    2- and 6-dim reference points"""
    if family == (12, 20):
        _needs_ab(msda)
    msda._lib.set_tuning("fwd_variant", family[0])
    msda._lib.set_tuning("bwd_variant", family[1])
    try:
        _fused_vs_unfused(msda, cuda_device, case, dtype, ref_dim)
    finally:
        msda._lib.set_tuning("bwd_variant", -1); msda._lib.set_tuning("fwd_variant", -1)


@pytest.mark.parametrize("family", [(-1, -1), (11, 21), (12, 20)], ids=["default", "binned", "tile"])
def test_fused_preprocessing_on_the_encoder_pyramid(msda, cuda_device, family):
    """the encoder's call: queries = the pixel pyramid, reference points = pixel centres"""
    from monosowa_b200.workloads import encoder_reference_points
    if family == (12, 20):
        _needs_ab(msda)
    shapes = [(24, 80), (12, 40), (6, 20), (3, 10)]
    S = sum(h * w for h, w in shapes)
    ref = encoder_reference_points(shapes, cuda_device)[None].expand(2, -1, -1, -1).contiguous()
    msda._lib.set_tuning("fwd_variant", family[0]); msda._lib.set_tuning("bwd_variant", family[1])
    try:
        for dtype in (torch.float32, torch.bfloat16):
            _fused_vs_unfused(msda, cuda_device, (shapes, 2, 8, 32, S, 4), dtype, 2, ref=ref, off_sigma=2.0)
    finally:
        msda._lib.set_tuning("fwd_variant", -1); msda._lib.set_tuning("bwd_variant", -1)


def _fused_vs_unfused(msda, cuda_device, case, dtype, ref_dim=2, ref=None, off_sigma=3.0):
    from monosowa_b200.ops.functions import MSDeformAttnFusedFunction, fused_supported
    from monosowa_b200.ops.modules.ms_deform_attn import sampling_locations_from_reference
    shapes, N, M, D, Lq, P = case
    dev = cuda_device
    g = torch.Generator().manual_seed(77 + D + Lq)
    sh, lsi = _levels(shapes)
    S, L = int(sh.prod(1).sum()), len(shapes)
    value = torch.randn(N, S, M, D, generator=g).to(dev, dtype)
    if ref is None:
        ref = (torch.rand(N, Lq, L, 2, generator=g) * 1.2 - 0.1).to(dev)
        if ref_dim == 6:                                      # (cx, cy, l, r, t, b): box extents of a few percent of the image
            ref = torch.cat([ref, (torch.rand(N, Lq, L, 4, generator=g) * 0.3 + 0.02).to(dev)], -1)
    offs = (torch.randn(N, Lq, M, L, P, 2, generator=g) * off_sigma).to(dev)
    logits = (torch.randn(N, Lq, M, L * P, generator=g) * 2.0).to(dev)
    grad_out = torch.randn(N, Lq, M * D, generator=g).to(dev, dtype)
    shd, lsid = sh.to(dev), lsi.to(dev)
    assert fused_supported(value, ref, offs, logits, shd, lsid, L, P)
    assert not fused_supported(value, ref.cpu(), offs, logits, shd, lsid, L, P)          # wrong device -> literal path
    assert not fused_supported(value, ref, offs, logits, shd.int(), lsid, L, P)           # int32 shapes -> literal path
    want_ref_grad = ref_dim == 2

    def unfused(v, o, lg, r):
        loc = sampling_locations_from_reference(r, o, shd, P)
        aw = torch.softmax(lg, -1).view(N, Lq, M, L, P)
        return msda.MSDeformAttnFunction.apply(v, shd, lsid, loc.contiguous(), aw.contiguous(), 64), loc, aw

    res = []
    for fused in (True, False):
        v = value.clone().requires_grad_(True); o = offs.clone().requires_grad_(True); lg = logits.clone().requires_grad_(True)
        r = ref.clone().requires_grad_(want_ref_grad)
        if fused:
            out = MSDeformAttnFusedFunction.apply(v, shd, lsid, r, o, lg)
        else:
            out, loc, aw = unfused(v, o, lg, r)
        out.backward(grad_out)
        res.append((out.detach(), v.grad, o.grad, lg.grad, r.grad))
    (fo, fgv, fgo, fgl, fgr), (uo, ugv, ugo, ugl, ugr) = res
    t_out, t_gv = (2e-6, 1e-5) if dtype == torch.float32 else (4e-3, 1e-2)
    assert fgv.dtype == dtype
    assert O.rel_l2(fo, uo) < t_out
    assert O.rel_l2(fgv, ugv) < t_gv
    assert O.rel_l2(fgl, ugl) < 1e-5
    # offsets: same floor() decisions (identical fp32 location arithmetic), so no masking needed
    assert O.rel_l2(fgo, ugo) < 1e-5
    if want_ref_grad:
        assert O.rel_l2(fgr, ugr) < 1e-5
    # and DIRECTLY against the fp64 oracle on the locations / weights torch produced (fp32-exact inputs)
    loc64, aw64 = loc.detach().cpu().double(), aw.detach().cpu().double()
    ref_out = O.forward_c(value.cpu().double(), sh, lsi, loc64, aw64)
    assert O.rel_l2(fo, ref_out) < (1e-5 if dtype == torch.float32 else 4e-3)
    rgv, rgl, rga = O.backward_c(value.cpu().double(), sh, lsi, loc64, aw64, grad_out.cpu().double())
    assert O.rel_l2(fgv, rgv) < (1e-4 if dtype == torch.float32 else 1e-2)
    # chain rule through the pre-processing in fp64: softmax backward, d loc / d offset
    rg_logit = (aw64 * (rga - (aw64 * rga).sum((-1, -2), keepdim=True))).reshape(N, Lq, M, L * P)
    assert O.rel_l2(fgl, rg_logit) < 1e-4
    if ref_dim == 2:
        scale = 1.0 / torch.stack([sh[:, 1], sh[:, 0]], -1).double()[None, None, None, :, None, :]
    else:
        r64 = ref.cpu().double()[:, :, None, :, None, :]
        scale = (r64[..., 2::2] + r64[..., 3::2]) * 0.5 / P
    keep = ~O.pixel_boundary_mask(loc.detach().cpu(), sh, eps_px=1e-4)
    assert O.rel_l2(fgo.cpu().double()[keep], (rgl * scale)[keep]) < 1e-4
    if want_ref_grad:
        rgl_m = torch.where(keep, rgl, torch.zeros_like(rgl))
        fgl_back = torch.where(keep, fgo.cpu().double() / scale, torch.zeros_like(rgl))
        assert O.rel_l2(fgl_back.sum((2, 4)), rgl_m.sum((2, 4))) < 1e-4


def test_module_takes_the_fused_path_for_every_call_of_a_training_forward(msda, cuda_device, monkeypatch):
    """All six MSDA calls of a MonoDETR training forward qualify: the encoder (2-dim pixel-centre references, no
    gradient), decoder layer 0 (2-dim learned references WITH a gradient, depthaware_transformer.py:286) and decoder
    layers 1-2 (6-dim detached boxes, :613).  6-dim references that need a gradient, or wrong devices, stay literal."""
    from monosowa_b200.ops.modules import ms_deform_attn as modfile
    calls = []
    real_f, real_u = modfile.MSDeformAttnFusedFunction.apply, modfile.MSDeformAttnFunction.apply

    class F_:
        apply = staticmethod(lambda *a: (calls.append("fused"), real_f(*a))[1])

    class U_:
        apply = staticmethod(lambda *a: (calls.append("unfused"), real_u(*a))[1])

    monkeypatch.setattr(modfile, "MSDeformAttnFusedFunction", F_)
    monkeypatch.setattr(modfile, "MSDeformAttnFunction", U_)
    dev = cuda_device
    sh, lsi = _levels([(6, 8), (3, 4)])
    mod = msda.MSDeformAttn(64, 2, 4, 2).to(dev)
    q, src = torch.randn(2, 7, 64, device=dev), torch.randn(2, 60, 64, device=dev)
    ref2 = torch.rand(2, 7, 2, 2, device=dev)
    ref6 = torch.cat([ref2, torch.rand(2, 7, 2, 4, device=dev) * 0.3], -1)
    out_f = mod(q, ref2, src, sh.to(dev), lsi.to(dev))
    mod.fuse_preprocessing = False
    out_u = mod(q, ref2, src, sh.to(dev), lsi.to(dev))
    out6_u = mod(q, ref6, src, sh.to(dev), lsi.to(dev))
    mod.fuse_preprocessing = True
    r2g = ref2.clone().requires_grad_(True)
    mod(q, r2g, src, sh.to(dev), lsi.to(dev)).sum().backward()                        # 2-dim refs with a gradient -> fused
    assert r2g.grad is not None and torch.isfinite(r2g.grad).all()
    out6_f = mod(q, ref6, src, sh.to(dev), lsi.to(dev))                               # 6-dim detached refs -> fused
    mod(q, ref6.clone().requires_grad_(True), src, sh.to(dev), lsi.to(dev))           # 6-dim refs with a gradient -> literal
    assert calls == ["fused", "unfused", "unfused", "fused", "fused", "unfused"]
    assert O.rel_l2(out_f, out_u) < 2e-6 and O.rel_l2(out6_f, out6_u) < 2e-6


# --------------------------------------------------------------------------------------------
# 7. the MSDeformAttn module against goldens produced by the reference module
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ref2", "ref6"])
def test_module_matches_reference_module_golden(msda, cuda_device, name):
    g = load_golden("module", name)
    mod = msda.MSDeformAttn(d_model=g["d_model"], n_levels=g["shapes"].shape[0], n_heads=g["heads"],
                            n_points=g["points"]).double()
    state = {k[len("state__"):]: v for k, v in g.items() if k.startswith("state__")}
    mod.load_state_dict(state, strict=True)
    mod = mod.to(cuda_device)
    dev = cuda_device
    out = mod(g["query"].to(dev), g["ref"].to(dev), g["src"].to(dev), g["shapes"].to(dev),
              g["level_start_index"].to(dev), g["mask"].to(dev))
    assert O.rel_l2(out, g["out"]) < 1e-12
    out32 = mod.float()(g["query"].float().to(dev), g["ref"].float().to(dev), g["src"].float().to(dev),
                        g["shapes"].to(dev), g["level_start_index"].to(dev), g["mask"].to(dev))
    assert O.rel_l2(out32, g["out"]) < 1e-5


def test_module_bf16_autocast_trains(msda, cuda_device):
    torch.manual_seed(0)
    dev = cuda_device
    sh, lsi = _levels([(12, 40), (6, 20), (3, 10), (2, 5)])
    S = int(sh.prod(1).sum())
    mod = msda.MSDeformAttn().to(dev)
    q = torch.randn(2, S, 256, device=dev, requires_grad=True)
    src = torch.randn(2, S, 256, device=dev, requires_grad=True)
    from monosowa_b200.workloads import encoder_reference_points
    ref = encoder_reference_points(sh.tolist(), dev)[None].expand(2, -1, -1, -1).contiguous()
    out32 = mod(q, ref, src, sh.to(dev), lsi.to(dev))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out16 = mod(q, ref, src, sh.to(dev), lsi.to(dev))
    assert out16.dtype == torch.bfloat16
    assert O.rel_l2(out16.float(), out32) < 2e-2
    out16.float().pow(2).mean().backward()
    assert q.grad is not None and src.grad is not None and torch.isfinite(src.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in mod.parameters())


# --------------------------------------------------------------------------------------------
# 8. the measurement build (libmsda_b200_ab.so, -DMSDA_AB): tile kernels, kept correct although not shipped
# --------------------------------------------------------------------------------------------
def test_measurement_build_suite_in_a_child_process(msda):
    """The tile kernels exist only in the measurement build, which is loaded instead of the product
    library when MSDA_AB=1.  A process has one library, so their cases -- skipped above -- run in a child process."""
    import os
    import subprocess
    import sys
    from monosowa_b200 import build as B
    if msda._lib.has_ab_flavours():
        pytest.skip("already running against the measurement build")
    if not os.path.exists(B.AB_LIB):
        pytest.skip("libmsda_b200_ab.so not built (python -m monosowa_b200.build --ab)")
    env = dict(os.environ, MSDA_AB="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "long_query or every_launch_variant or fused_preprocessing or guard_bands or nan_behind"],
                       capture_output=True, text=True, timeout=1500, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in tail and "skipped" not in tail, tail

