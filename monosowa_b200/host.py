"""Host-buffer step: MSDA forward + backward for tensors that live in (pinned) HOST memory.

Thin binding of ``msda_host_step_{f32,bf16}`` (include/msda_b200.h): the C library pipelines the batch
image by image -- H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1 overlap -- so a caller
with host data pays roughly the PCIe time of the larger direction instead of copy + kernels + copy.
The reference extension has no counterpart (CUDA tensors only, ops/src/ms_deform_attn.h:29-38).

    out, grad_value, grad_loc, grad_attn = host_step(value, shapes_dev, lsi_dev, loc, attn, grad_out)

``value`` / ``loc`` / ``attn`` / ``grad_out`` are CPU tensors (pin them for asynchronous copies);
``shapes_dev`` / ``lsi_dev`` are int64 CUDA tensors and select the device.  Results are pinned CPU tensors
(pass ``results=`` to reuse buffers); ``grad_value`` has ``value``'s dtype.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_SUFFIX = {torch.float32: "f32", torch.bfloat16: "bf16"}
_workspaces: dict = {}           # (device, stream) -> staging buffer: concurrent streams never share stages


def _workspace(dev: torch.device, stream: int, nbytes: int) -> torch.Tensor:
    ws = _workspaces.get((dev, stream))
    if ws is None or ws.numel() < nbytes:
        ws = _workspaces[(dev, stream)] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    return ws


def release_workspaces() -> None:
    """drop the cached device staging buffers (one per device, 3 pipeline stages of images_per_chunk images each)"""
    _workspaces.clear()


def host_step(value, spatial_shapes, level_start_index, sampling_locations, attention_weights, grad_output,
              images_per_chunk: int = 1, results=None, synchronize: bool = True, stages: int = 3):
    dev = spatial_shapes.device
    if dev.type != "cuda" or level_start_index.device != dev:
        raise NotImplementedError("spatial_shapes / level_start_index must be CUDA tensors: they select the device "
                                  "(MSDA has no CPU implementation)")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    host = dict(value=value, sampling_loc=sampling_locations, attn_weight=attention_weights, grad_output=grad_output)
    for name, t in host.items():
        if t.is_cuda:
            raise RuntimeError(f"{name} is a CUDA tensor: use MSDeformAttnFunction for device-resident data")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if value.dtype not in _SUFFIX or grad_output.dtype != value.dtype:
        raise RuntimeError(f"unsupported dtypes value {value.dtype} / grad_output {grad_output.dtype} (float32 or bfloat16)")
    if sampling_locations.dtype != torch.float32 or attention_weights.dtype != torch.float32:
        raise RuntimeError("sampling_loc and attn_weight must be float32")
    if value.dim() != 4 or sampling_locations.dim() != 6 or attention_weights.dim() != 5:
        raise RuntimeError("expected value (N,S,M,D), sampling_loc (N,Lq,M,L,P,2), attn_weight (N,Lq,M,L,P)")
    n, s, m, d = value.shape
    _, lq, m2, nl, p, two = sampling_locations.shape
    if (sampling_locations.shape[0] != n or m2 != m or two != 2 or tuple(attention_weights.shape) != (n, lq, m, nl, p)
            or tuple(spatial_shapes.shape) != (nl, 2) or level_start_index.numel() != nl
            or grad_output.numel() != n * lq * m * d):
        raise RuntimeError("inconsistent shapes")
    if results is None:
        results = (torch.empty((n, lq, m * d), dtype=value.dtype).pin_memory(),
                   torch.empty(value.shape, dtype=value.dtype).pin_memory(),
                   torch.empty(sampling_locations.shape, dtype=torch.float32).pin_memory(),
                   torch.empty(attention_weights.shape, dtype=torch.float32).pin_memory())
    out, gv, gl, ga = results
    for name, t, numel, dt in (("out", out, n * lq * m * d, value.dtype), ("grad_value", gv, value.numel(), value.dtype),
                               ("grad_loc", gl, sampling_locations.numel(), torch.float32),
                               ("grad_attn", ga, attention_weights.numel(), torch.float32)):
        if t.is_cuda or not t.is_contiguous() or t.numel() != numel or t.dtype != dt:
            raise RuntimeError(f"result buffer {name}: expected a contiguous CPU tensor of {numel} {dt} elements, "
                               f"got {tuple(t.shape)} {t.dtype} on {t.device}")
    is_bf16 = int(value.dtype == torch.bfloat16)
    need = int(_lib.lib.msda_host_step_workspace_bytes(is_bf16, s, m, d, nl, lq, p, int(images_per_chunk)))
    fn = getattr(_lib.lib, "msda_host_step_" + _SUFFIX[value.dtype])
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream()
        ws_bytes = max(need // 3 * max(3, min(int(stages), 16)), 256)          # `need` = the minimum: 3 stages
        ws = _workspace(dev, stream.cuda_stream, ws_bytes)
        rc = fn(ptr(value), ptr(spatial_shapes), ptr(level_start_index), ptr(sampling_locations), ptr(attention_weights),
                ptr(grad_output), ptr(out), ptr(gv), ptr(gl), ptr(ga), ptr(ws), ctypes.c_size_t(ws_bytes),
                n, s, m, d, nl, lq, p, int(images_per_chunk), ctypes.c_void_p(stream.cuda_stream))
        if rc:
            raise RuntimeError(f"msda_host_step failed (code {rc}): {_lib.last_error()}")
        if synchronize:
            stream.synchronize()
    return out, gv, gl, ga
