"""the UNMODIFIED reference MonoDETR (staged under baseline/_ref by tools/stage_reference.py)
takes one deterministic training step with this repo's op swapped in, and the same step with the reference's own
ops/ Python on the reference's own CUDA kernels recompiled for sm_100a (oracle/_ref).  Loss terms and a sample of
parameter gradients must agree.

Tolerances (fp32; BASELINE.json north_star for the op: forward <= 1e-5, backward <= 1e-4 relative, the latter
allowing for atomic ordering): total loss and every loss term rel <= 1e-5; sampled gradients of the transformer's
parameters (everything downstream of the op: the MSDA modules' own Linears included) rel-L2 <= 2e-5 each; sampled
gradients of the ResNet-50 backbone rel-L2 <= 5e-4 -- those sit behind ~50 convolution layers whose cuDNN
backward-filter kernels and this op's value-gradient atomics both sum in a run-dependent order: TWO RUNS OF THE SAME
ARM differ by 1.0-1.6e-4 there (measured on B200 and printed by the test as `self-noise`), so 1e-4 is below the
noise floor of the comparison itself.  Each arm runs in its own process: the op is swapped through sys.modules at
import time, so the two cannot share an interpreter.

History: this test is what exposed that the reference BINARY computes the pixel coordinate with one fused
multiply-add (msda_common.cuh make_tap) -- at initialisation every sample of the model sits on a pixel boundary and
the sampling_offsets bias gradient of the encoder differed by 2.6e-3 until the kernels did the same."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "MonoDETR")
pytestmark = pytest.mark.gpu


def _dump(tmp_path, op, host_opt="none", batch=2, extra=()):
    out = tmp_path / f"step_{op}_{host_opt}{'_'.join(extra)}.pt"
    cmd = [sys.executable, os.path.join(ROOT, "tools", "train_step_bench.py"), "--op", op, "--host-opt", host_opt,
           "--batch", str(batch), "--dump-step", str(out), *extra]
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, f"{' '.join(cmd)} failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}"
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["op"] == op
    return torch.load(out)


def _rel(a, b):
    return abs(a - b) / max(abs(b), 1e-12)


TOL_GRAD = {"depthaware_transformer": 2e-5, "": 5e-4}      # by parameter-name prefix, see the module docstring


def _tol(name, tol_grad):
    if not isinstance(tol_grad, dict):
        return tol_grad
    return next(v for k, v in tol_grad.items() if name.startswith(k))


def _compare(got, want, tol_loss, tol_grad, label):
    assert set(got["loss_terms"]) == set(want["loss_terms"]) and len(want["loss_terms"]) >= 30
    assert _rel(got["loss"], want["loss"]) <= tol_loss, f"{label}: loss {got['loss']} vs {want['loss']}"
    worst_term = max(_rel(got["loss_terms"][k], want["loss_terms"][k]) for k in want["loss_terms"])
    assert worst_term <= tol_loss, f"{label}: a loss term differs by {worst_term}"
    assert set(got["grads"]) == set(want["grads"]) and len(want["grads"]) >= 40
    # Some parameters have a mathematically ZERO gradient (a bias added to every key of a softmax attention shifts all
    # logits alike): what both arms hold there is rounding noise ~1e-7 of the global norm, so the denominator is floored.
    floor = 1e-6 * want["grad_norm_all"]
    errs = {}
    for name, g in want["grads"].items():
        den = max(g.double().norm().item(), floor)
        errs[name] = (got["grads"][name].double() - g.double()).norm().item() / den
    worst = max(errs, key=errs.get)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"{label}: largest gradient differences {[(n, float(f'{e:.2e}')) for n, e in top]}")
    print(f"{label}: loss rel {_rel(got['loss'], want['loss']):.2e}, worst loss term {worst_term:.2e}, "
          f"{len(errs)} sampled gradients, worst rel-L2 {errs[worst]:.2e} ({worst}), "
          f"global grad norm {got['grad_norm_all']:.6g} vs {want['grad_norm_all']:.6g}")
    for name, e in errs.items():
        assert e <= _tol(name, tol_grad), f"{label}: gradient of {name} differs by rel-L2 {e} (tolerance {_tol(name, tol_grad)})"
    assert _rel(got["grad_norm_all"], want["grad_norm_all"]) <= 1e-5


@pytest.fixture(scope="module")
def reference_arm(tmp_path_factory):
    from oracle import msda_oracle as O
    if not os.path.isdir(REF):
        pytest.skip("baseline/_ref/MonoDETR not staged (python tools/stage_reference.py where /root/reference exists)")
    if not O.ref_cuda_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return _dump(tmp_path_factory.mktemp("ref"), "ref_cuda")


def test_training_step_with_our_op_matches_the_reference_kernels(reference_arm, tmp_path):
    """this repo's module (all six MSDA calls on the fused path) + kernels vs the reference's module + kernels"""
    ours = _dump(tmp_path, "ours")
    _compare(ours, reference_arm, 1e-5, TOL_GRAD, "ours vs ref_cuda")
    again = _dump(tmp_path, "ours")                       # the comparison's own noise floor: the same arm twice
    noise = {n: (again["grads"][n].double() - g.double()).norm().item() / max(g.double().norm().item(), 1e-6 * ours["grad_norm_all"])
             for n, g in ours["grads"].items()}
    worst = max(noise, key=noise.get)
    print(f"self-noise (ours vs ours, two runs): worst sampled gradient rel-L2 {noise[worst]:.2e} ({worst})")
    literal = _dump(tmp_path, "ours", extra=("--no-fuse",))
    _compare(literal, reference_arm, 1e-5, TOL_GRAD, "ours, literal (unfused) module path vs ref_cuda")


def test_training_step_with_device_resident_host_sections_matches(reference_arm, tmp_path):
    """SURVEY 8 f3 on top: matcher, DDN target painting and AdamW replaced by their device-resident versions
    (monosowa_b200.step_host) -- the step's loss terms and gradients must not move"""
    ours = _dump(tmp_path, "ours", host_opt="all")
    _compare(ours, reference_arm, 1e-5, TOL_GRAD, "ours + step_host vs ref_cuda")
