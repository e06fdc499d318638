"""CPU oracle for the MSDA hot path.  TEST INFRASTRUCTURE ONLY -- never the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  ``monosowa_b200`` never does: the
product path has no CPU fallback and fails loudly without its CUDA library.

Two independent restatements of the reference algorithm live here, so that they can be
checked against each other as well as against the golden vectors:

* :func:`core_grid_sample` -- the reference's own test oracle, restated:
  ``ms_deform_attn_core_pytorch`` (MonoDETR/lib/models/monodetr/ops/functions/
  ms_deform_attn_func.py:41-61): per level, reshape the value slab to ``(N*M, D, H, W)``,
  map locations to ``[-1, 1]`` and ``F.grid_sample(bilinear, zeros, align_corners=False)``,
  then the attention-weighted sum over levels x points.  Its arithmetic lives in PyTorch's
  ``grid_sampler_2d`` CPU kernel (torch 2.11.0 here; the reference pins torch 1.13.1,
  MonoDETR/requirements.txt:2).  Gradients come from autograd.
* :func:`forward_c` / :func:`backward_c` -- ``oracle/msda_oracle.c``: explicit per-sample
  bilinear loops following the CUDA kernels (ms_deform_im2col_cuda.cuh:33-159, 237-299,
  347-401) with analytic gradients, in fp64 or in reference-faithful fp32.

Parity pin: both are checked against ``tests/golden/*.npz`` (outputs of the *reference's
own* ``ms_deform_attn_core_pytorch`` imported from /root/reference by
``tests/golden/gen_golden.py``) in ``tests/test_oracle.py``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmsda_oracle.so")
_REF_LIB_PATH = os.path.join(_HERE, "_ref", "libmsda_ref_sm100.so")
_lib = None
_ref_lib = None


# ----------------------------------------------------------------------------------------
# restatement 1: grid_sample path (ms_deform_attn_func.py:41-61)
# ----------------------------------------------------------------------------------------
def core_grid_sample(value, spatial_shapes, sampling_locations, attention_weights):
    """value (N,S,M,D); spatial_shapes (L,2) [(H,W)]; loc (N,Lq,M,L,P,2) in (x,y) order,
    normalised; attention_weights (N,Lq,M,L,P).  Returns (N, Lq, M*D)."""
    n, s, m, d = value.shape
    lq, nl, npts = sampling_locations.shape[1], sampling_locations.shape[3], sampling_locations.shape[4]
    hw = [(int(h), int(w)) for h, w in spatial_shapes.tolist()]
    assert sum(h * w for h, w in hw) == s and len(hw) == nl
    grids = sampling_locations * 2 - 1                       # func.py:48
    per_level = []
    start = 0
    for lvl, (h, w) in enumerate(hw):
        slab = value[:, start:start + h * w]                 # (N, HW, M, D)        func.py:47
        start += h * w
        img = slab.permute(0, 2, 3, 1).reshape(n * m, d, h, w)                     # func.py:52
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(n * m, lq, npts, 2)  # :54
        per_level.append(F.grid_sample(img, grid, mode="bilinear", padding_mode="zeros",
                                       align_corners=False))                        # :56-57
    sampled = torch.stack(per_level, dim=3).reshape(n * m, d, lq, nl * npts)
    wts = attention_weights.permute(0, 2, 1, 3, 4).reshape(n * m, 1, lq, nl * npts)  # :60
    out = (sampled * wts).sum(-1).reshape(n, m * d, lq)                            # :61
    return out.transpose(1, 2).contiguous()


def core_grid_sample_fwd_bwd(value, spatial_shapes, loc, attn, grad_out):
    """Forward + autograd backward through :func:`core_grid_sample`.
    Returns (out, grad_value, grad_loc, grad_attn)."""
    v = value.detach().clone().requires_grad_(True)
    l = loc.detach().clone().requires_grad_(True)
    a = attn.detach().clone().requires_grad_(True)
    out = core_grid_sample(v, spatial_shapes, l, a)
    out.backward(grad_out.reshape(out.shape))
    return out.detach(), v.grad, l.grad, a.grad


# ----------------------------------------------------------------------------------------
# restatement 2: explicit loops in C
# ----------------------------------------------------------------------------------------
def build_c(force: bool = False) -> str:
    """Compile oracle/msda_oracle.c (and, when /root/reference is present, oracle/_ref)."""
    src = os.path.join(_HERE, "msda_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "libmsda_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_c()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _np(t, dtype):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=dtype)


def _dims(value, loc):
    n, s, m, d = value.shape
    lq, nl, npts = loc.shape[1], loc.shape[3], loc.shape[4]
    return n, s, m, d, nl, lq, npts


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward_c(value, spatial_shapes, level_start_index, loc, attn, precision="f64"):
    """Explicit-loop oracle forward.  Returns a torch tensor (N, Lq, M*D) of the chosen precision."""
    dt = np.float64 if precision == "f64" else np.float32
    v, l, a = _np(value, dt), _np(loc, dt), _np(attn, dt)
    sh, st = _np(spatial_shapes, np.int64), _np(level_start_index, np.int64)
    n, s, m, d, nl, lq, npts = _dims(v, l)
    out = np.empty((n, lq, m * d), dtype=dt)
    fn = getattr(_load(), f"msda_oracle_forward_{precision}")
    fn.restype = None
    fn(_ptr(v), _ptr(sh), _ptr(st), _ptr(l), _ptr(a), _ptr(out),
       *(ctypes.c_int(x) for x in (n, s, m, d, nl, lq, npts)))
    return torch.from_numpy(out)


def backward_c(value, spatial_shapes, level_start_index, loc, attn, grad_out, precision="f64"):
    """Explicit-loop oracle backward.  Returns (grad_value, grad_loc, grad_attn)."""
    dt = np.float64 if precision == "f64" else np.float32
    v, l, a, g = _np(value, dt), _np(loc, dt), _np(attn, dt), _np(grad_out, dt)
    sh, st = _np(spatial_shapes, np.int64), _np(level_start_index, np.int64)
    n, s, m, d, nl, lq, npts = _dims(v, l)
    gv = np.zeros_like(v)
    gl = np.empty_like(l)
    ga = np.empty_like(a)
    fn = getattr(_load(), f"msda_oracle_backward_{precision}")
    fn.restype = None
    fn(_ptr(v), _ptr(sh), _ptr(st), _ptr(l), _ptr(a), _ptr(g), _ptr(gv), _ptr(gl), _ptr(ga),
       *(ctypes.c_int(x) for x in (n, s, m, d, nl, lq, npts)))
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(ga)


# ----------------------------------------------------------------------------------------
# the reference's own CUDA kernels (oracle/_ref), GPU baseline + second checker
# ----------------------------------------------------------------------------------------
def ref_cuda_available() -> bool:
    return os.path.exists(_REF_LIB_PATH)


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        _ref_lib = ctypes.CDLL(_REF_LIB_PATH)
    return _ref_lib


def _sfx(t):
    return {torch.float32: "f32", torch.float64: "f64"}[t.dtype]


def ref_cuda_forward(value, spatial_shapes, level_start_index, loc, attn):
    """Run the reference's ms_deformable_im2col_cuda (cuh:923-954) on CUDA tensors."""
    n, s, m, d, nl, lq, npts = _dims(value, loc)
    out = torch.empty((n, lq, m * d), dtype=value.dtype, device=value.device)
    fn = getattr(_load_ref(), f"msda_ref_forward_{_sfx(value)}")
    rc = fn(*(ctypes.c_void_p(t.data_ptr()) for t in (value, spatial_shapes, level_start_index, loc, attn, out)),
            *(ctypes.c_int(x) for x in (n, s, m, d, nl, lq, npts)),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(f"reference forward kernel failed: cudaError {rc}")
    return out


def ref_cuda_backward(value, spatial_shapes, level_start_index, loc, attn, grad_out):
    """Run the reference's ms_deformable_col2im_cuda (cuh:956-1327) on CUDA tensors."""
    n, s, m, d, nl, lq, npts = _dims(value, loc)
    gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
    fn = getattr(_load_ref(), f"msda_ref_backward_{_sfx(value)}")
    rc = fn(*(ctypes.c_void_p(t.data_ptr()) for t in
              (value, spatial_shapes, level_start_index, loc, attn, grad_out, gv, gl, ga)),
            *(ctypes.c_int(x) for x in (n, s, m, d, nl, lq, npts)),
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(f"reference backward kernel failed: cudaError {rc}")
    return gv, gl, ga


# ----------------------------------------------------------------------------------------
# error metrics shared by the parity tests
# ----------------------------------------------------------------------------------------
def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    den = b.norm().item()
    return (a - b).norm().item() / den if den > 0 else (a - b).norm().item()


def max_abs_over_max(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / den if den > 0 else (a - b).abs().max().item()


def pixel_boundary_mask(loc, spatial_shapes, eps_px=1e-4):
    """True where a sample's pixel coordinate is within ``eps_px`` of an integer in either
    axis -- there d(out)/d(loc) is discontinuous and floor() may land on either side
    depending on the arithmetic precision.  loc (N,Lq,M,L,P,2) -> bool mask
    of the same shape (both components of a flagged sample are masked)."""
    sh = spatial_shapes.to(loc.device).double()
    wh = torch.stack([sh[:, 1], sh[:, 0]], -1)[None, None, None, :, None, :]
    px = loc.double() * wh - 0.5
    near = (px - px.round()).abs() < eps_px
    flag = near[..., 0] | near[..., 1]
    return flag[..., None].expand_as(loc)
