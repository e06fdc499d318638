"""FrozenBatchNorm2d (+ residual add) (+ ReLU) of the MonoDETR backbone in one pass -- bit-identical to the reference.

The reference normalises with ``x * scale + bias`` on broadcast tensors (MonoDETR/lib/models/monodetr/backbone.py:55-65;
five tiny kernels to form scale / bias from the four frozen buffers on EVERY call, then two element-wise passes over the
activation) and torchvision's ``Bottleneck.forward`` follows with ``out += identity`` and an in-place ReLU.  At the KITTI
input an activation of ``layer1`` is 503 MB (16 x 256 x 96 x 320 fp32): each pass is ~0.15 ms, there are 53
normalisations, and autograd mirrors every pass.  ``fuse_frozen_bn(model)`` patches the live modules (no reference file
is edited) so that

* ``FrozenBatchNorm2d.forward``                    -> one kernel  (``detr_frozen_bn_act_f32``),
* ``Bottleneck.forward``: bn1 + relu, bn2 + relu  -> one kernel each,
                          bn3 + identity + relu   -> one kernel,

with scale / bias formed ONCE by the reference's own expression and cached until a buffer changes.  Every arithmetic
step keeps its own rounding in the reference's order, so outputs and gradients are bitwise equal
(tests/test_step_host.py::test_fused_frozen_bn_is_bitwise_identical).  Anything the kernels do not cover (CPU tensors,
other dtypes -- bf16 autocast --, non-contiguous or non-4-D inputs) takes the reference's own code.
"""
from __future__ import annotations

import ctypes
import types

import torch
from torch.autograd import Function

from . import lsa as _lsa


def _lib():
    lib = _lsa.lib()
    if not getattr(lib, "_bn_bound", False):
        vp, ll, it = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
        lib.detr_frozen_bn_act_f32.restype = it
        lib.detr_frozen_bn_act_f32.argtypes = [vp, vp, vp, vp, vp, ll, ll, it, it, vp]
        lib.detr_frozen_bn_act_backward_f32.restype = it
        lib.detr_frozen_bn_act_backward_f32.argtypes = [vp, vp, vp, vp, vp, ll, ll, it, it, vp]
        lib._bn_bound = True
    return lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _covered(x, residual=None):
    ok = x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.is_contiguous()
    if ok and residual is not None:
        ok = residual.shape == x.shape and residual.dtype == x.dtype and residual.device == x.device and residual.is_contiguous()
    return ok


class _FrozenBNAct(Function):
    @staticmethod
    def forward(ctx, x, scale, bias, residual, relu):
        y = torch.empty_like(x)
        n, c, h, w = x.shape
        with torch.cuda.device(x.device):
            rc = _lib().detr_frozen_bn_act_f32(_p(x), _p(residual), _p(scale), _p(bias), _p(y), x.numel(), h * w, c, int(relu),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc:
            raise RuntimeError(f"detr_frozen_bn_act_f32 failed (code {rc})")
        ctx.relu, ctx.has_res = bool(relu), residual is not None
        ctx.save_for_backward(scale, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        scale, y = ctx.saved_tensors
        grad_y = grad_y.contiguous()
        need_x, need_res = ctx.needs_input_grad[0], ctx.has_res and ctx.needs_input_grad[3]
        n, c, h, w = grad_y.shape
        gx = torch.empty_like(grad_y)
        # the identity branch receives the masked gradient itself; without a ReLU that IS grad_y (no copy needed)
        gres = torch.empty_like(grad_y) if (need_res and ctx.relu) else None
        if need_x or gres is not None:
            with torch.cuda.device(grad_y.device):
                rc = _lib().detr_frozen_bn_act_backward_f32(_p(grad_y), _p(y), _p(scale), _p(gx), _p(gres), grad_y.numel(), h * w, c,
                                                            int(ctx.relu), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc:
                raise RuntimeError(f"detr_frozen_bn_act_backward_f32 failed (code {rc})")
        if need_res and not ctx.relu:
            gres = grad_y
        return (gx if need_x else None), None, None, (gres if need_res else None), None


def _scale_bias(bn):
    """scale / bias exactly as backbone.py:58-64 forms them, cached while the four frozen buffers are unchanged"""
    key = tuple((t._version, t.data_ptr(), t.device, t.dtype) for t in (bn.weight, bn.bias, bn.running_var, bn.running_mean))
    cached = getattr(bn, "_msda_scale_bias", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            w, b = bn.weight.reshape(1, -1, 1, 1), bn.bias.reshape(1, -1, 1, 1)
            rv, rm = bn.running_var.reshape(1, -1, 1, 1), bn.running_mean.reshape(1, -1, 1, 1)
            scale = w * (rv + bn.eps).rsqrt()
            bias = b - rm * scale
        cached = (key, scale.reshape(-1).contiguous(), bias.reshape(-1).contiguous())
        bn._msda_scale_bias = cached
    return cached[1], cached[2]


def frozen_bn_act(bn, x, residual=None, relu=False):
    """relu?(bn(x) + residual?) through the fused kernel; `bn` is a live FrozenBatchNorm2d"""
    scale, bias = _scale_bias(bn)
    return _FrozenBNAct.apply(x, scale, bias, residual, relu)


def _is_frozen_bn(m):
    return type(m).__name__ == "FrozenBatchNorm2d" and all(hasattr(m, a) for a in ("weight", "bias", "running_var", "running_mean", "eps"))


def _bn_forward(self, x):
    if _covered(x) and self.weight.dtype == torch.float32:
        return frozen_bn_act(self, x)
    return self._msda_reference_forward(x)


def _bottleneck_forward(self, x):
    """torchvision.models.resnet.Bottleneck.forward with its three normalisations fused with what follows them"""
    out = self.conv1(x)
    if not (_covered(out) and self.bn1.weight.dtype == torch.float32):
        return self._msda_reference_forward(x)
    out = frozen_bn_act(self.bn1, out, relu=True)
    out = frozen_bn_act(self.bn2, self.conv2(out), relu=True)
    out = self.conv3(out)
    identity = x if self.downsample is None else self.downsample(x)
    if not _covered(out, identity):
        out = self.bn3(out)
        out += identity
        return self.relu(out)
    return frozen_bn_act(self.bn3, out, residual=identity.contiguous(), relu=True)


def fuse_frozen_bn(model):
    """Patch every FrozenBatchNorm2d and every torchvision Bottleneck built on them inside `model`; returns the number of
    modules patched (idempotent)."""
    n = 0
    for m in model.modules():
        if getattr(m, "_msda_fused_bn", False):
            continue
        if _is_frozen_bn(m):
            m._msda_reference_forward = m.forward
            m.forward = types.MethodType(_bn_forward, m)
            m._msda_fused_bn = True
            n += 1
        elif type(m).__name__ == "Bottleneck" and all(_is_frozen_bn(getattr(m, a, None)) for a in ("bn1", "bn2", "bn3")) \
                and isinstance(getattr(m, "relu", None), torch.nn.ReLU):
            m._msda_reference_forward = m.forward
            m.forward = types.MethodType(_bottleneck_forward, m)
            m._msda_fused_bn = True
            n += 1
    return n
