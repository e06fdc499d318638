"""CPU: host-side logic, the C-ABI surface, and the N>1 aggregation path (gloo, world size 2).
No kernel is launched here (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    import monosowa_b200
    from monosowa_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "msda_b200.h")).read()
    declared = set(re.findall(r"\b(msda_[a-z0-9_]+)\s*\(", hdr))
    assert {"msda_forward_f32", "msda_backward_f32", "msda_forward_bf16", "msda_backward_bf16",
            "msda_forward_f64", "msda_backward_f64", "msda_last_error", "msda_abi_version"} <= declared
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/msda_b200.h but not exported"
    assert set(_lib.EXPORTS) == declared
    assert _lib.lib.msda_abi_version() == _lib.ABI_VERSION == 2
    assert "sm_100a" in _lib.build_info()
    assert monosowa_b200.MSDeformAttnFunction is not None


def test_library_contains_sm100a_code_and_vector_reductions():
    """cuobjdump: the .so carries sm_100a SASS with 128-bit loads and REDG.128 (no PTX-only JIT path)."""
    from monosowa_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "RED.E.ADD.F32x4" in sass.replace("REDG", "RED") or "REDG.E.ADD.F32x4" in sass
    assert "LDG.E.128" in sass


def test_argument_errors_are_reported_without_a_gpu():
    from monosowa_b200 import _lib
    lib = _lib.lib
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    assert lib.msda_forward_f32(null, one, one, one, one, one, 1, 4, 1, 32, 1, 1, 4, null) == -1
    assert "value is NULL" in _lib.last_error()
    assert lib.msda_forward_f32(one, one, one, one, one, one, 1, 4, 1, 32, 17, 1, 4, null) == -2
    assert "MSDA_MAX_LEVELS" in _lib.last_error()
    assert lib.msda_backward_f32(one, one, one, one, one, one, one, one, one, 1, -4, 1, 32, 1, 1, 4, null) == -2
    assert lib.msda_forward_f32(ctypes.c_void_p(18), one, one, one, one, one, 1, 4, 1, 32, 1, 1, 4, null) == -3
    # empty problems succeed without touching the device
    assert lib.msda_forward_f32(null, one, one, null, null, null, 0, 4, 8, 32, 1, 5, 4, null) == 0
    assert _lib.last_error() == ""
    with pytest.raises(ValueError):
        _lib.set_tuning("no_such_knob", 1)
    _lib.set_tuning("fwd_pipe", 6)
    assert _lib.get_tuning("fwd_pipe") == 6
    _lib.set_tuning("fwd_pipe", -1)
    assert lib.msda_describe_forward(32, 0, 16, 8, 32, 4, 4, 550) == b"fwd_rec_f32"
    assert lib.msda_describe_backward(32, 0, 16, 8, 32, 4, 4, 550) == b"bwd_rec_f32"
    assert lib.msda_describe_backward(32, 0, 16, 8, 32, 4, 3, 50) == b"bwd_rec_f32"
    assert lib.msda_describe_backward(32, 1, 16, 8, 32, 4, 4, 50) == b"bwd_rec_bf16"
    assert lib.msda_describe_forward(64, 0, 16, 8, 32, 4, 4, 10200) == b"fwd_generic_f64"
    assert lib.msda_describe_backward(32, 0, 16, 8, 30, 4, 4, 10200) == b"bwd_generic_f32"
    # long query sets (the encoder): the backward combines the coarse levels in shared memory (binned kernel) when
    # there are enough (image, head, chunk) work items; the forward stays with the record kernel
    assert lib.msda_describe_forward(32, 0, 16, 8, 32, 4, 4, 10200) == b"fwd_rec_f32"
    assert lib.msda_describe_backward(32, 0, 16, 8, 32, 4, 4, 10200) == b"bwd_bin_f32"
    assert lib.msda_describe_backward(32, 1, 16, 8, 32, 4, 4, 10200) == b"bwd_bin_bf16"
    assert lib.msda_describe_backward(32, 0, 2, 8, 32, 4, 4, 10200) == b"bwd_rec_f32"       # too few work items
    assert _lib.describe("backward", torch.bfloat16, 16, 8, 32, 4, 4, 550) == "bwd_rec_bf16"
    # bf16 backward: short query sets scatter straight into the bf16 gradient, long ones need the fp32 scratch
    assert lib.msda_backward_bf16_scratch_bytes(16, 10200, 8, 32, 4, 550, 4, 1) == 0
    assert lib.msda_backward_bf16_scratch_bytes(1, 640, 8, 32, 4, 300, 4, 1) == 640 * 8 * 32 * 4      # dense: 30 additions per row
    assert lib.msda_backward_bf16_scratch_bytes(2, 10200, 8, 32, 4, 10200, 4, 1) == 2 * 10200 * 8 * 32 * 4
    assert lib.msda_backward_bf16_scratch_bytes(2, 100, 8, 30, 4, 50, 4, 1) == 2 * 100 * 8 * 30 * 4      # generic kernel
    assert lib.msda_backward_bf16_scratch_bytes(2, 100, 8, 32, 4, 50, 4, 0) == 2 * 100 * 8 * 32 * 4      # misaligned


def test_cpu_tensors_raise_not_implemented_no_fallback():
    import monosowa_b200 as msda
    v = torch.randn(1, 4, 2, 32)
    sh = torch.tensor([[2, 2]]); lsi = torch.tensor([0])
    loc = torch.rand(1, 3, 2, 1, 4, 2); aw = torch.rand(1, 3, 2, 1, 4)
    with pytest.raises(NotImplementedError):
        msda.MSDeformAttnFunction.apply(v, sh, lsi, loc, aw, 64)
    with pytest.raises(NotImplementedError):
        msda.MSDeformAttn(64, 1, 2, 4)(torch.randn(1, 3, 64), torch.rand(1, 3, 1, 2), torch.randn(1, 4, 64), sh, lsi)


def test_op_schema_and_meta_shapes():
    import monosowa_b200  # noqa: F401
    v = torch.empty(2, 30, 8, 32, device="meta")
    sh = torch.empty(2, 2, dtype=torch.long, device="meta"); lsi = torch.empty(2, dtype=torch.long, device="meta")
    loc = torch.empty(2, 7, 8, 2, 4, 2, device="meta"); aw = torch.empty(2, 7, 8, 2, 4, device="meta")
    out = torch.ops.msda.forward(v, sh, lsi, loc, aw, 64)
    assert out.shape == (2, 7, 256) and out.device.type == "meta"
    gv, gl, ga = torch.ops.msda.backward(v, sh, lsi, loc, aw, out, 64)
    assert gv.shape == v.shape and gl.shape == loc.shape and ga.shape == aw.shape


def test_module_parameters_match_reference_state_dict():
    import monosowa_b200 as msda
    g = load_golden("module", "ref2")
    ref_state = {k[len("state__"):]: v for k, v in g.items() if k.startswith("state__")}
    mod = msda.MSDeformAttn(d_model=g["d_model"], n_levels=2, n_heads=g["heads"], n_points=g["points"])
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == {k: tuple(v.shape) for k, v in ref_state.items()}
    mod.double().load_state_dict(ref_state, strict=True)
    m = msda.MSDeformAttn()
    assert m.im2col_step == 64 and (m.d_model, m.n_levels, m.n_heads, m.n_points) == (256, 4, 8, 4)
    # reference init (ms_deform_attn.py:106-120): zero weights, ring-of-directions bias
    assert m.sampling_offsets.weight.abs().max() == 0 and m.attention_weights.bias.abs().max() == 0
    b = m.sampling_offsets.bias.view(8, 4, 4, 2)
    assert torch.allclose(b[0, 0, :, 0], torch.tensor([1., 2., 3., 4.])) and torch.allclose(b[0, :, :, 1], torch.zeros(4, 4), atol=1e-6)
    assert torch.allclose(b[2, 1, 2], torch.tensor([0., 3.]), atol=1e-5)
    cross = msda.MSDeformAttn_cross(256, 4, 8, 4)
    assert cross.value_proj.weight.shape == (128, 128) and cross.conditional
    assert msda.MultiheadAttention is torch.nn.MultiheadAttention


@pytest.mark.parametrize("name", ["ref2", "ref6"])
def test_module_host_arithmetic_matches_reference_module(monkeypatch, name):
    """Module-level arithmetic (projections, softmax, 2-/6-dim reference points) against the golden
    produced by the reference module, with the op itself replaced by the oracle (CPU, test only)."""
    import monosowa_b200 as msda
    from monosowa_b200.ops.modules import ms_deform_attn as modfile
    from oracle import msda_oracle as O
    g = load_golden("module", name)

    class OracleFn:
        @staticmethod
        def apply(value, shapes, lsi, loc, aw, step):
            return O.core_grid_sample(value, shapes, loc, aw)

    monkeypatch.setattr(modfile, "MSDeformAttnFunction", OracleFn)
    mod = msda.MSDeformAttn(d_model=g["d_model"], n_levels=2, n_heads=g["heads"], n_points=g["points"]).double()
    mod.load_state_dict({k[len("state__"):]: v for k, v in g.items() if k.startswith("state__")})
    out = mod(g["query"], g["ref"], g["src"], g["shapes"], g["level_start_index"], g["mask"])
    assert O.rel_l2(out, g["out"]) < 1e-13
    with pytest.raises(ValueError):
        mod(g["query"], torch.rand(*g["ref"].shape[:3], 4, dtype=torch.float64), g["src"], g["shapes"], g["level_start_index"], g["mask"])


def test_workloads_and_byte_model():
    from monosowa_b200 import workloads as W
    assert W.KITTI == [(48, 160), (24, 80), (12, 40), (6, 20)]
    assert [sum(h * w for h, w in s) for s in (W.KITTI, W.KITTI360, W.WAYMO, W.ALT640)] == [10200, 11044, 51000, 12750]
    ab = W.algorithmic_bytes(W.config(1))
    assert round(ab["fwd"] / 1e6, 1) == 584.9 and round(ab["bwd"] / 1e6, 1) == 1002.7   # SURVEY.md 8d worked values
    assert round(W.algorithmic_bytes(W.config(0))["total"] / 1e6, 1) == 198.5
    d = W.make_inputs(W.config(0, batch=1))
    assert d["loc"].shape == (1, 10200, 8, 4, 4, 2) and d["lsi"].tolist() == [0, 7680, 9600, 10080]
    oob = ((d["loc"] < 0) | (d["loc"] > 1)).any(-1).float().mean().item()
    assert 0.10 < oob < 0.20                                   # 15-18 % in the hooked model
    assert torch.allclose(d["attn"].sum((-1, -2)), torch.ones(1, 10200, 8), atol=1e-5)
    d2 = W.make_inputs(W.config(2, batch=1))
    assert d2["value"].dtype == torch.bfloat16 and d2["loc"].dtype == torch.float32 and d2["loc"].shape[1] == 50


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench

    def reduce_fn(t, op):
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t

    # rank 1 is slower: the job is as fast as its slowest rank, bytes add up
    val, ms = bench.aggregate(1.0e9, 2.0 + rank, world, reduce_fn)
    q.put((rank, val, ms))
    dist.destroy_process_group()


def test_multi_rank_aggregation_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _, val, ms in res:
        assert ms == 3.0 and abs(val - 2.0e9 / 3.0e-3 / 1e9) < 1e-9


def test_bench_reference_arm_contract():
    """--impl reference prints one JSON line with the agreed keys (bounded CPU sample)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    staged = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MonoDETR", "lib", "models", "monodetr", "ops",
                                         "functions", "ms_deform_attn_func.py"))
    # the reference's own function when the tree is staged (baseline/_ref), the oracle's restatement of it otherwise
    assert line["cpu_baseline"]["kind"] == ("reference" if staged else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"].startswith("BASELINE.json configs[1]")
    assert line["native_libraries_loaded"] == [], "the reference arm must not load the product library"


def test_reference_arm_function_is_the_oracles_restatement():
    """bench.py --impl reference times the staged reference's ms_deform_attn_core_pytorch; the oracle restates the same
    function: on the same inputs they must agree bit for bit (both are grid_sample on this torch build)"""
    import bench
    from oracle import msda_oracle as O
    core = bench.load_reference_core()
    if core is None:
        pytest.skip("baseline/_ref/MonoDETR not staged")
    g = torch.Generator().manual_seed(3)
    sh = torch.tensor([[6, 4], [3, 2]])
    value = torch.rand(2, 30, 2, 8, generator=g)
    loc = torch.rand(2, 5, 2, 2, 3, 2, generator=g) * 1.4 - 0.2
    attn = torch.softmax(torch.rand(2, 5, 2, 6, generator=g), -1).view(2, 5, 2, 2, 3)
    assert torch.equal(core(value, sh, loc, attn), O.core_grid_sample(value, sh, loc, attn))
