"""Autograd boundary of the MSDA op -- drop-in for the reference's
MonoDETR/lib/models/monodetr/ops/functions/ms_deform_attn_func.py:21-38.

``MSDeformAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
sampling_locations, attention_weights, im2col_step)`` keeps the reference signature, return
shape ``(N, Lq, M*D)`` and gradient tuple ``(grad_value, None, None, grad_sampling_loc,
grad_attn_weight, None)``.  Underneath, instead of the pybind11 extension
``MultiScaleDeformableAttention`` (src/vision.cpp:13-16), two ``torch.library`` ops
``msda::forward`` / ``msda::backward`` call the C ABI of libmsda_b200.so through ctypes.

Error behaviour mirrors the reference host wrapper (src/cuda/ms_deform_attn_cuda.cu:28-52):
RuntimeError for non-contiguous inputs and for ``batch % min(batch, im2col_step) != 0``;
CPU tensors raise NotImplementedError (the reference: AT_ERROR("Not implemented on the CPU"),
src/ms_deform_attn.h:38) -- there is no CPU fallback by design.
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib

_SUFFIX = {torch.float32: "f32", torch.float64: "f64", torch.bfloat16: "bf16"}
_LIBDEF = torch.library.Library("msda", "DEF")
_LIBDEF.define("forward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
               "Tensor sampling_locations, Tensor attention_weights, int im2col_step) -> Tensor")
_LIBDEF.define("backward(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
               "Tensor sampling_locations, Tensor attention_weights, Tensor grad_output, "
               "int im2col_step) -> (Tensor, Tensor, Tensor)")
_LIBDEF.define("forward_fused(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
               "Tensor reference_points, Tensor sampling_offsets, Tensor attention_logits) -> Tensor")
_LIBDEF.define("backward_fused(Tensor value, Tensor spatial_shapes, Tensor level_start_index, "
               "Tensor reference_points, Tensor sampling_offsets, Tensor attention_logits, "
               "Tensor grad_output, bool need_ref_grad) -> (Tensor, Tensor, Tensor, Tensor)")


def _check(value, spatial_shapes, level_start_index, loc, attn, im2col_step, grad_output=None):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", loc), ("attn_weight", attn)]
    if grad_output is not None:
        named.append(("grad_output", grad_output))
    for name, t in named:
        if not t.is_cuda:
            raise NotImplementedError(f"{name} must be a CUDA tensor: MSDA is not implemented on the CPU "
                                      "(no fallback; see oracle/ for the test-only CPU restatement)")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value is on {value.device}")
    if value.dtype not in _SUFFIX:
        raise RuntimeError(f"unsupported value dtype {value.dtype} (float32, float64, bfloat16)")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    if value.dim() != 4 or loc.dim() != 6 or attn.dim() != 5 or spatial_shapes.dim() != 2:
        raise RuntimeError("expected value (N,S,M,D), sampling_loc (N,Lq,M,L,P,2), attn_weight (N,Lq,M,L,P), "
                           "spatial_shapes (L,2)")
    n, s, m, d = value.shape
    _, lq, m2, nl, p, two = loc.shape
    if (loc.shape[0] != n or m2 != m or two != 2 or tuple(attn.shape) != (n, lq, m, nl, p)
            or spatial_shapes.shape[0] != nl or spatial_shapes.shape[1] != 2 or level_start_index.numel() != nl):
        raise RuntimeError(f"inconsistent shapes: value {tuple(value.shape)}, sampling_loc {tuple(loc.shape)}, "
                           f"attn_weight {tuple(attn.shape)}, spatial_shapes {tuple(spatial_shapes.shape)}")
    step = min(n, int(im2col_step))
    if n > 0 and (step <= 0 or n % step != 0):                                  # ms_deform_attn_cuda.cu:50-52
        raise RuntimeError(f"batch({n}) must divide im2col_step({step})")
    return n, s, m, d, nl, lq, p


def _coord_dtype(value):
    return torch.float64 if value.dtype == torch.float64 else torch.float32


def _as_coord(t, value):
    want = _coord_dtype(value)
    return t if t.dtype == want else t.to(want)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class _on_device:
    """Make `dev` current for the launch; free when it already is (the common case)."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(self.idx))

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def _raise(rc, what):
    raise RuntimeError(f"{what} failed (code {rc}): {_lib.last_error()}")


def _forward_cuda(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
    n, s, m, d, nl, lq, p = _check(value, spatial_shapes, level_start_index, sampling_locations,
                                   attention_weights, im2col_step)
    loc, attn = _as_coord(sampling_locations, value), _as_coord(attention_weights, value)
    out = torch.empty((n, lq, m * d), dtype=value.dtype, device=value.device)
    fn = getattr(_lib.lib, "msda_forward_" + _SUFFIX[value.dtype])
    with _on_device(value.device) as stream:
        rc = fn(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(loc), _ptr(attn), _ptr(out),
                n, s, m, d, nl, lq, p, stream)
    if rc:
        _raise(rc, "msda::forward")
    return out


def _aligned16(*tensors):
    return int(all(t.data_ptr() % 16 == 0 for t in tensors))


def _bf16_scratch(value, dims, *tensors):
    """fp32 accumulation buffer for the bf16 backward, when the library asks for one (long query sets)."""
    n, s, m, d, nl, lq, p = dims
    nbytes = _lib.lib.msda_backward_bf16_scratch_bytes(n, s, m, d, nl, lq, p, _aligned16(value, *tensors))
    return torch.empty(nbytes // 4, dtype=torch.float32, device=value.device) if nbytes else None


def _backward_cuda(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                   grad_output, im2col_step):
    grad_output = grad_output.contiguous()
    dims = _check(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step,
                  grad_output)
    n, s, m, d, nl, lq, p = dims
    if grad_output.numel() != n * lq * m * d:
        raise RuntimeError(f"grad_output has {grad_output.numel()} elements, expected {n * lq * m * d}")
    if grad_output.dtype != value.dtype:
        grad_output = grad_output.to(value.dtype)
    loc, attn = _as_coord(sampling_locations, value), _as_coord(attention_weights, value)
    ct = _coord_dtype(value)
    grad_value = torch.empty_like(value)                       # zero-filled by the library; bf16 for bf16 values
    grad_loc = torch.empty(loc.shape, dtype=ct, device=value.device)
    grad_attn = torch.empty(attn.shape, dtype=ct, device=value.device)
    fn = getattr(_lib.lib, "msda_backward_" + _SUFFIX[value.dtype])
    args = [_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(loc), _ptr(attn),
            _ptr(grad_output), _ptr(grad_value), _ptr(grad_loc), _ptr(grad_attn)]
    if value.dtype == torch.bfloat16:
        scratch = _bf16_scratch(value, dims, loc, attn, grad_output, grad_value, grad_loc, grad_attn)
        args.append(_ptr(scratch) if scratch is not None else None)
    with _on_device(value.device) as stream:
        rc = fn(*args, n, s, m, d, nl, lq, p, stream)
    if rc:
        _raise(rc, "msda::backward")
    if grad_loc.dtype != sampling_locations.dtype:
        grad_loc = grad_loc.to(sampling_locations.dtype)
    if grad_attn.dtype != attention_weights.dtype:
        grad_attn = grad_attn.to(attention_weights.dtype)
    return grad_value, grad_loc, grad_attn


# ------------------------------------------------------------------------------------------------
# fused pre-processing: softmax over L*P and the reference-point arithmetic
# (reference ops/modules/ms_deform_attn.py:145-155, 2- and 6-dim reference points) inside the kernels
# ------------------------------------------------------------------------------------------------
def fused_supported(value, reference_points, sampling_offsets, attention_logits, spatial_shapes, level_start_index,
                    n_levels, n_points):
    """True when the fused kernels can take this call (else the module uses the unfused op, whose own checks
    raise the reference's errors): everything on value's CUDA device, int64 shapes, D in {16,32,64},
    L*P <= D, 2- or 6-dim reference points."""
    d = value.shape[-1]
    if not (value.is_cuda and value.dtype in (torch.float32, torch.bfloat16) and d in (16, 32, 64)
            and n_levels * n_points <= d and reference_points.shape[-1] in (2, 6)):
        return False
    for t in (reference_points, sampling_offsets, attention_logits, spatial_shapes, level_start_index):
        if not t.is_cuda or t.device != value.device:
            return False
    return (spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64
            and spatial_shapes.is_contiguous() and level_start_index.is_contiguous() and value.is_contiguous())


def _fused_views(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits,
                 grad_output=None):
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("reference_points", reference_points), ("sampling_offsets", sampling_offsets),
             ("attention_logits", attention_logits)]
    if grad_output is not None:
        named.append(("grad_output", grad_output))
    for name, t in named:
        if not t.is_cuda:
            raise NotImplementedError(f"{name} must be a CUDA tensor: MSDA is not implemented on the CPU")
        if t.device != value.device:
            raise RuntimeError(f"{name} is on {t.device}, value is on {value.device}")
    for name, t in named[:3]:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    if value.dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"unsupported value dtype {value.dtype} for the fused op (float32, bfloat16)")
    if value.dim() != 4 or sampling_offsets.dim() != 6 or reference_points.dim() != 4:
        raise RuntimeError("expected value (N,S,M,D), reference_points (N,Lq,L,2|6), sampling_offsets (N,Lq,M,L,P,2)")
    n, s, m, d = value.shape
    _, lq, _, nl, p, _ = sampling_offsets.shape
    rd = reference_points.shape[-1]
    if tuple(reference_points.shape) != (n, lq, nl, rd) or rd not in (2, 6) \
            or attention_logits.numel() != n * lq * m * nl * p or sampling_offsets.shape[0] != n \
            or sampling_offsets.shape[2] != m or sampling_offsets.shape[5] != 2 \
            or spatial_shapes.shape[0] != nl or level_start_index.numel() != nl:
        raise RuntimeError(f"inconsistent shapes for the fused op: value {tuple(value.shape)}, "
                           f"ref {tuple(reference_points.shape)}, offsets {tuple(sampling_offsets.shape)}, "
                           f"logits {tuple(attention_logits.shape)}")
    f32 = lambda t: (t if t.dtype == torch.float32 else t.float()).contiguous()
    return (n, s, m, d, nl, lq, p), rd, f32(reference_points), f32(sampling_offsets), f32(attention_logits)


def _forward_fused_cuda(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
    dims, rd, ref, off, logit = _fused_views(value, spatial_shapes, level_start_index, reference_points,
                                             sampling_offsets, attention_logits)
    n, s, m, d, nl, lq, p = dims
    out = torch.empty((n, lq, m * d), dtype=value.dtype, device=value.device)
    fn = getattr(_lib.lib, "msda_forward_fused_" + _SUFFIX[value.dtype])
    with _on_device(value.device) as stream:
        rc = fn(_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(ref), rd, _ptr(off), _ptr(logit),
                _ptr(out), n, s, m, d, nl, lq, p, stream)
    if rc:
        _raise(rc, "msda::forward_fused")
    return out


def _backward_fused_cuda(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                         attention_logits, grad_output, need_ref_grad):
    grad_output = grad_output.contiguous()
    dims, rd, ref, off, logit = _fused_views(value, spatial_shapes, level_start_index, reference_points,
                                             sampling_offsets, attention_logits, grad_output)
    n, s, m, d, nl, lq, p = dims
    if grad_output.numel() != n * lq * m * d:
        raise RuntimeError(f"grad_output has {grad_output.numel()} elements, expected {n * lq * m * d}")
    if grad_output.dtype != value.dtype:
        grad_output = grad_output.to(value.dtype)
    grad_value = torch.empty_like(value)
    grad_off = torch.empty(off.shape, dtype=torch.float32, device=value.device)
    grad_logit = torch.empty(logit.shape, dtype=torch.float32, device=value.device)
    fn = getattr(_lib.lib, "msda_backward_fused_" + _SUFFIX[value.dtype])
    args = [_ptr(value), _ptr(spatial_shapes), _ptr(level_start_index), _ptr(ref), rd, _ptr(off), _ptr(logit),
            _ptr(grad_output), _ptr(grad_value), _ptr(grad_off), _ptr(grad_logit)]
    if value.dtype == torch.bfloat16:
        scratch = _bf16_scratch(value, dims, ref, off, logit, grad_output, grad_value, grad_off, grad_logit)
        args.append(_ptr(scratch) if scratch is not None else None)
    with _on_device(value.device) as stream:
        rc = fn(*args, n, s, m, d, nl, lq, p, stream)
    if rc:
        _raise(rc, "msda::backward_fused")
    grad_ref = _reference_point_grad(ref, grad_off, spatial_shapes, p) if need_ref_grad else grad_off.new_empty(0)
    return (grad_value, grad_off.to(sampling_offsets.dtype), grad_logit.to(attention_logits.dtype),
            grad_ref.to(reference_points.dtype))


def _reference_point_grad(ref, grad_off, spatial_shapes, n_points):
    """d loss / d reference_points from d loss / d sampling_offsets (the chain rule of
    ops/modules/ms_deform_attn.py:149-155, summed over heads and points)."""
    if ref.shape[-1] == 2:                       # loc = ref + off / (W, H): d loc / d ref = 1, grad_loc = grad_off * (W, H)
        wh = torch.stack([spatial_shapes[:, 1], spatial_shapes[:, 0]], -1).to(grad_off.dtype)
        return (grad_off * wh[None, None, None, :, None, :]).sum((2, 4))
    # loc = ref[:2] + off / P * ext * 0.5 with ext = (r2 + r3, r4 + r5):
    #   grad_loc = grad_off * P / (0.5 * ext);  d loc / d r[2..5] = off / P * 0.5
    raise NotImplementedError("gradient w.r.t. 6-dim reference points: use the unfused path")


_LIBIMPL = torch.library.Library("msda", "IMPL")
_LIBIMPL.impl("forward", _forward_cuda, "CUDA")
_LIBIMPL.impl("backward", _backward_cuda, "CUDA")
_LIBIMPL.impl("forward_fused", _forward_fused_cuda, "CUDA")
_LIBIMPL.impl("backward_fused", _backward_fused_cuda, "CUDA")


@torch.library.register_fake("msda::forward_fused")
def _forward_fused_fake(value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
    n, _, m, d = value.shape
    return value.new_empty((n, sampling_offsets.shape[1], m * d))


@torch.library.register_fake("msda::backward_fused")
def _backward_fused_fake(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                         attention_logits, grad_output, need_ref_grad):
    return (torch.empty_like(value), torch.empty_like(sampling_offsets), torch.empty_like(attention_logits),
            torch.empty_like(reference_points) if need_ref_grad else reference_points.new_empty(0))


@torch.library.register_fake("msda::forward")
def _forward_fake(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, im2col_step):
    n, _, m, d = value.shape
    return value.new_empty((n, sampling_locations.shape[1], m * d))


@torch.library.register_fake("msda::backward")
def _backward_fake(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                   grad_output, im2col_step):
    return (torch.empty_like(value), torch.empty_like(sampling_locations), torch.empty_like(attention_weights))


def _setup_context(ctx, inputs, output):
    value, shapes, lsi, loc, attn, step = inputs
    ctx.im2col_step = step
    ctx.save_for_backward(value, shapes, lsi, loc, attn)


def _autograd_backward(ctx, grad_output):
    value, shapes, lsi, loc, attn = ctx.saved_tensors
    gv, gl, ga = torch.ops.msda.backward(value, shapes, lsi, loc, attn, grad_output, ctx.im2col_step)
    return gv, None, None, gl, ga, None


torch.library.register_autograd("msda::forward", _autograd_backward, setup_context=_setup_context)


class MSDeformAttnFunction(Function):
    """Same contract as the reference class (ms_deform_attn_func.py:21-38)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = torch.ops.msda.forward(value, value_spatial_shapes, value_level_start_index,
                                        sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = torch.ops.msda.backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output, ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


class MSDeformAttnFusedFunction(Function):
    """MSDA with the module's pre-processing folded into the kernels:
    ``apply(value, spatial_shapes, level_start_index, reference_points (N,Lq,L,2|6),
    sampling_offsets (N,Lq,M,L,P,2), attention_logits (N,Lq,M,L*P))`` equals
    ``MSDeformAttnFunction.apply(value, shapes, lsi, locations, softmax(logits, -1).view(N,Lq,M,L,P), im2col_step)``
    with ``locations`` as the reference module forms them (ops/modules/ms_deform_attn.py:149-155).
    2-dim ``reference_points`` are differentiated (the decoder's first layer feeds learned ones); 6-dim ones
    must not require a gradient (the reference detaches them, depthaware_transformer.py:613)."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, reference_points, sampling_offsets, attention_logits):
        out = torch.ops.msda.forward_fused(value, spatial_shapes, level_start_index, reference_points,
                                           sampling_offsets, attention_logits)
        ctx.save_for_backward(value, spatial_shapes, level_start_index, reference_points, sampling_offsets,
                              attention_logits)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        need_ref = ctx.needs_input_grad[3]
        gv, goff, glogit, gref = torch.ops.msda.backward_fused(*ctx.saved_tensors, grad_output, need_ref)
        return gv, None, None, (gref if need_ref else None), goff, glogit


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """Name-compatible with the pybind export ``MSDA.ms_deform_attn_forward`` (src/vision.cpp:14)."""
    return torch.ops.msda.forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """Name-compatible with ``MSDA.ms_deform_attn_backward`` (src/vision.cpp:15); returns a list of 3."""
    return list(torch.ops.msda.backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                        grad_output, im2col_step))
