"""Build libmsda_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m monosowa_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the repo snapshot.  Replaces the reference's ops/setup.py + make.sh (which target sm_60-75
through torch's CUDAExtension and refuse to build without a visible GPU, setup.py:36-52).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(PKG, "libmsda_b200.so")
SOURCES = ["msda_forward.cu", "msda_backward.cu", "msda_backward_binned.cu", "msda_backward_tiled.cu", "msda_capi.cu"]
HEADERS = [os.path.join(CSRC, "msda_common.cuh"), os.path.join(CSRC, "msda_records.cuh"), os.path.join(CSRC, "msda_tiles.cuh"),
           os.path.join(ROOT, "include", "msda_b200.h"),
           os.path.join(ROOT, "include", "monodetr_step_b200.h")]
# second library: device-side pieces of the MonoDETR training step, kept out of the operator's ABI
STEP_LIB = os.path.join(PKG, "libmonodetr_step_b200.so")
STEP_SOURCES = ["step_lsa.cu", "step_bn.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmsda_b200.so cannot be built (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc: str, src: str, verbose: bool, extra=(), tag="") -> str:
    obj = os.path.join(BUILD, os.path.splitext(src)[0] + tag + ".o")
    srcp = os.path.join(CSRC, src)
    if _stale(obj, [srcp, *HEADERS, os.path.abspath(__file__)]):
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", srcp, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
    return obj


def _build_one(lib: str, sources, force: bool, verbose: bool, extra=(), tag="") -> str:
    srcs = [os.path.join(CSRC, s) for s in sources]
    if not force and not _stale(lib, [*srcs, *HEADERS, os.path.abspath(__file__)]):
        return lib
    nvcc = nvcc_path()
    os.makedirs(BUILD, exist_ok=True)
    if force:
        for s in sources:
            for f in (os.path.splitext(s)[0] + tag + ".o", os.path.splitext(s)[0] + tag + ".o.log"):
                if os.path.exists(os.path.join(BUILD, f)):
                    os.remove(os.path.join(BUILD, f))
    with cf.ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, verbose, extra, tag), sources))
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    """libmsda_b200.so (the operator) -- returns its path; also keeps libmonodetr_step_b200.so up to date."""
    _build_one(STEP_LIB, STEP_SOURCES, force, verbose)
    return _build_one(LIB, SOURCES, force, verbose)


def build_step(force: bool = False, verbose: bool = False) -> str:
    return _build_one(STEP_LIB, STEP_SOURCES, force, verbose)


AB_LIB = os.path.join(PKG, "libmsda_b200_ab.so")


def build_ab(force: bool = False, verbose: bool = False) -> str:
    """The measurement build: the same sources with -DMSDA_AB, i.e. with the extra launch flavours that
    msda_set_tuning("fwd_pipe" / "bwd_pipe") selects.  A separate, git-ignored file; the product library never
    carries them.  Loaded instead of libmsda_b200.so when the environment has MSDA_AB=1 (tools/sweep.py)."""
    return _build_one(AB_LIB, SOURCES, force, verbose, extra=("-DMSDA_AB",), tag=".ab")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--ab", action="store_true", help="also build libmsda_b200_ab.so (A/B launch flavours)")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
    if a.ab:
        print(build_ab(a.force, a.verbose))
