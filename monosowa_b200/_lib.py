"""ctypes loader for libmsda_b200.so (C ABI in include/msda_b200.h).

There is deliberately no fallback: if the CUDA library is missing and cannot be built,
importing this module raises.  The product never routes through oracle/ or any CPU path.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmsda_b200.so")
if os.environ.get("MSDA_AB") == "1":                     # measurement build with the A/B launch flavours (build.py --ab)
    LIB_PATH = os.path.join(_PKG, "libmsda_b200_ab.so")

_vp, _i64p, _int = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int
_FWD = [_vp, _i64p, _i64p, _vp, _vp, _vp] + [_int] * 7 + [_vp]
_BWD = [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp]
_BWD_BF16 = [_vp, _i64p, _i64p, _vp, _vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp]       # + scratch_f32
_FFWD = [_vp, _i64p, _i64p, _vp, _int, _vp, _vp, _vp] + [_int] * 7 + [_vp]                     # ref, ref_dim, ...
_FBWD = [_vp, _i64p, _i64p, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp]
_FBWD_BF16 = [_vp, _i64p, _i64p, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp] + [_int] * 7 + [_vp]   # + scratch_f32

_HOST = [_vp] * 11 + [ctypes.c_size_t] + [_int] * 8 + [_vp]

EXPORTS = {
    "msda_forward_f32": (_int, _FWD), "msda_forward_f64": (_int, _FWD), "msda_forward_bf16": (_int, _FWD),
    "msda_backward_f32": (_int, _BWD), "msda_backward_f64": (_int, _BWD), "msda_backward_bf16": (_int, _BWD_BF16),
    "msda_backward_bf16_scratch_bytes": (ctypes.c_size_t, [_int] * 8),
    "msda_forward_fused_f32": (_int, _FFWD), "msda_forward_fused_bf16": (_int, _FFWD),
    "msda_backward_fused_f32": (_int, _FBWD), "msda_backward_fused_bf16": (_int, _FBWD_BF16),
    "msda_host_step_f32": (_int, _HOST), "msda_host_step_bf16": (_int, _HOST),
    "msda_host_step_workspace_bytes": (ctypes.c_size_t, [_int] * 8),
    "msda_abi_version": (_int, []),
    "msda_build_info": (ctypes.c_char_p, []),
    "msda_last_error": (ctypes.c_char_p, []),
    "msda_launch_count": (ctypes.c_longlong, []),
    "msda_set_tuning": (_int, [ctypes.c_char_p, _int]),
    "msda_get_tuning": (_int, [ctypes.c_char_p]),
    "msda_describe_forward": (ctypes.c_char_p, [_int] * 8),
    "msda_describe_backward": (ctypes.c_char_p, [_int] * 8),
}
ABI_VERSION = 2
ERR_UNSUPPORTED = -4


def _bind(lib: ctypes.CDLL) -> ctypes.CDLL:
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch
        fn.restype, fn.argtypes = res, args
    if lib.msda_abi_version() != ABI_VERSION:
        raise AttributeError(f"ABI version {lib.msda_abi_version()} != expected {ABI_VERSION}")
    return lib


def _open() -> ctypes.CDLL:
    from . import build as _build
    if not os.path.exists(LIB_PATH):
        try:
            _build.build_ab() if LIB_PATH.endswith("_ab.so") else _build.build()
        except Exception as exc:  # noqa: BLE001
            raise ImportError(
                f"monosowa_b200: {LIB_PATH} is missing and could not be built ({exc}). "
                "Run `python monosowa_b200/build.py`. There is no CPU or PyTorch fallback.") from exc
    lib = ctypes.CDLL(LIB_PATH)
    try:
        return _bind(lib)
    except AttributeError as stale:
        # a library left over from an older source tree (a symbol of include/msda_b200.h is missing):
        # rebuild once from the sources next to it, then fail loudly if it still does not match
        import _ctypes
        _ctypes.dlclose(lib._handle)
        try:
            _build.build(force=True)
            return _bind(ctypes.CDLL(LIB_PATH))
        except Exception as exc:  # noqa: BLE001
            raise ImportError(f"monosowa_b200: {LIB_PATH} does not match include/msda_b200.h ({stale}) and the "
                              f"rebuild failed ({exc}). There is no CPU or PyTorch fallback.") from exc


lib = _open()


def last_error() -> str:
    return lib.msda_last_error().decode()


def launch_count() -> int:
    return int(lib.msda_launch_count())


def set_tuning(key: str, value: int) -> None:
    if lib.msda_set_tuning(key.encode(), int(value)) != 0:
        raise ValueError(last_error())


def get_tuning(key: str) -> int:
    return int(lib.msda_get_tuning(key.encode()))


def build_info() -> str:
    return lib.msda_build_info().decode()


def describe(direction: str, dtype, N: int, M: int, D: int, L: int, P: int, Lq: int) -> str:
    """kernel family the library picks: describe("backward", torch.float32, 16, 8, 32, 4, 4, 10200) -> "bwd_bin_f32"."""
    name = str(dtype)
    bits, bf = (64, 0) if "64" in name else (32, 1 if "bfloat16" in name else 0)
    fn = lib.msda_describe_forward if direction == "forward" else lib.msda_describe_backward
    return fn(bits, bf, N, M, D, L, P, Lq).decode()


def has_ab_flavours() -> bool:
    """True for the measurement build (libmsda_b200_ab.so, MSDA_AB=1): the tile kernels exist there."""
    return "+AB" in build_info()
