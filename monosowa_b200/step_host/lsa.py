"""ctypes binding of libmonodetr_step_b200.so (include/monodetr_step_b200.h)."""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "libmonodetr_step_b200.so")
ERR_UNSUPPORTED = -4
MAX_IMAGES = 256

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            from .. import build as _build
            _build.build_step()
        handle = ctypes.CDLL(LIB_PATH)
        handle.detr_group_lsa_f32.restype = ctypes.c_int
        handle.detr_group_lsa_f32.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3
        handle.detr_group_lsa_status_f32.restype = ctypes.c_int
        handle.detr_group_lsa_status_f32.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 4
        handle.detr_step_last_error.restype = ctypes.c_char_p
        _lib = handle
    return _lib


class _Status:
    """Non-finite-cost flag of the matching kernel, read WITHOUT adding a synchronisation to the step: the device flag is
    copied to pinned host memory behind the kernel and looked at when the next call arrives (by then the copy of the
    previous call has long completed).  scipy.optimize.linear_sum_assignment raises ValueError on NaN / Inf costs
    (reference matcher.py:101); so does this, one matcher call late."""

    def __init__(self, device):
        self.dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.event = None

    def check(self, wait=False):
        if self.event is not None and (wait or self.event.query()):
            if wait:
                self.event.synchronize()
            self.event = None
            if int(self.host[0]) != 0:
                self.host.zero_()
                self.dev.zero_()
                raise ValueError("matrix contains invalid numeric entries (NaN or Inf in the matching cost of an earlier "
                                 "call: the device matcher reports it when it is next used)")

    def arm(self):
        self.host.copy_(self.dev, non_blocking=True)
        self.event = torch.cuda.Event()
        self.event.record()


_status: dict = {}


def check_status(device=None, wait=True):
    """raise now if a matching call on `device` (default: every device used so far) saw non-finite costs"""
    for dev, st in _status.items():
        if device is None or torch.device(device) == dev:
            st.check(wait=wait)


def group_lsa(cost: torch.Tensor, sizes, groups: int):
    """cost (B, Q, T) float32 CUDA, sizes = targets per image (host ints, sum == T).  Returns per-image lists of
    (query_index, target_index) int64 CUDA tensors in the order matcher.py:99-104 produces (group after group,
    ascending query inside a group), or None when a sub-problem exceeds the kernel's shared memory."""
    if not cost.is_cuda or cost.dtype != torch.float32 or cost.dim() != 3:
        raise RuntimeError("cost must be a (B, Q, T) float32 CUDA tensor")
    cost = cost.contiguous()
    B, Q, T = cost.shape
    sizes = [int(s) for s in sizes]
    if len(sizes) != B or sum(sizes) != T or Q % groups != 0:
        raise RuntimeError(f"inconsistent sizes: cost {tuple(cost.shape)}, sizes {sizes}, groups {groups}")
    if B > MAX_IMAGES:
        return None
    nq = Q // groups
    per_image = [groups * min(s, nq) for s in sizes]
    total = sum(per_image)
    out_q = torch.empty(total, dtype=torch.int64, device=cost.device)
    out_t = torch.empty(total, dtype=torch.int64, device=cost.device)
    if total:
        with torch.cuda.device(cost.device):
            st = _status.get(cost.device)
            if st is None:
                st = _status[cost.device] = _Status(cost.device)
            st.check()                                                       # an earlier call's verdict, if it has arrived
            rc = lib().detr_group_lsa_status_f32(ctypes.c_void_p(cost.data_ptr()), (ctypes.c_int * B)(*sizes), B, Q, T, groups,
                                                 ctypes.c_void_p(out_q.data_ptr()), ctypes.c_void_p(out_t.data_ptr()),
                                                 ctypes.c_void_p(st.dev.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if rc == 0:
                st.arm()
        if rc == ERR_UNSUPPORTED:
            return None
        if rc:
            raise RuntimeError(f"detr_group_lsa_f32 failed (code {rc}): {lib().detr_step_last_error().decode()}")
    pairs = MatchList(zip(out_q.split(per_image), out_t.split(per_image)))
    pairs.query_flat, pairs.target_flat, pairs.per_image = out_q, out_t, per_image
    return pairs


class MatchList(list):
    """the reference's list of per-image (query_index, target_index) tuples, plus the flat device tensors the kernel
    wrote them into (the concatenations the criterion rebuilds for every loss term)"""
    query_flat = target_flat = per_image = None
