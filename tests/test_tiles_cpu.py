"""Host-side check of the query-tile plan (monosowa_b200/csrc/msda_tiles.cuh): the per-level rectangles of the tiles
must PARTITION every level's pixels -- each query of a pixel-pyramid query set is then processed exactly once by the
tile kernels of the measurement build, whatever the pyramid looks like.  `tile_lo` is a __host__ __device__ function,
so the very code the kernels run is compiled for the host here (nvcc, no GPU needed)."""
import ctypes
import os
import random
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "msda_tiles.cuh"
extern "C" int msda_test_tile_lo(int t, int n_t, int tdim, int size, int bsize) { return msda::tile_lo(t, n_t, tdim, size, bsize); }
extern "C" int msda_test_floor_div(long a, long b) { return msda::floor_div(a, b); }
extern "C" int msda_test_tile_h(void) { return msda::kTileH; }
extern "C" int msda_test_tile_w(void) { return msda::kTileW; }
'''


@pytest.fixture(scope="module")
def tiles(tmp_path_factory):
    from monosowa_b200 import build as B
    d = tmp_path_factory.mktemp("tiles")
    src, lib = d / "tiles_host.cu", d / "libtiles_host.so"
    src.write_text(SRC)
    cmd = [B.nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-I" + os.path.join(ROOT, "include"), "-I" + B.CSRC, "-o", str(lib), str(src)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    h = ctypes.CDLL(str(lib))
    h.msda_test_floor_div.argtypes = [ctypes.c_long, ctypes.c_long]
    return h


def test_floor_div_rounds_towards_minus_infinity(tiles):
    for a, b in ((7, 2), (-7, 2), (-8, 2), (0, 5), (-1, 1000), (999, 1000), (-1000, 1000), (-1001, 1000)):
        assert tiles.msda_test_floor_div(a, b) == a // b


def test_tile_rectangles_partition_every_level(tiles):
    rng = random.Random(7)
    th, tw = tiles.msda_test_tile_h(), tiles.msda_test_tile_w()
    cases = [(48, 48), (48, 24), (48, 12), (48, 6), (47, 24), (160, 20), (13, 7), (1, 1), (5, 16384), (16384, 3)]
    cases += [(rng.randint(1, 400), rng.randint(1, 400)) for _ in range(300)]
    for tdim in (th, tw):
        for bsize, size in cases:                       # bsize: extent of the tiling base level, size: of some level
            n_t = (bsize + tdim - 1) // tdim
            lo = [tiles.msda_test_tile_lo(t, n_t, tdim, size, bsize) for t in range(n_t + 1)]
            assert lo[0] == 0 and lo[-1] == size, (bsize, size, lo)
            assert all(a <= b for a, b in zip(lo, lo[1:])), (bsize, size, lo)          # disjoint, ordered, complete
            for t in range(n_t):
                for y in range(lo[t], lo[t + 1]):       # every pixel sits in the tile that contains its centre
                    centre_in_base = (2 * y + 1) * bsize                                 # (y + 0.5) / size * bsize, times 2 * size
                    assert 2 * size * t * tdim <= centre_in_base or t == 0, (bsize, size, t, y)
                    assert centre_in_base < 2 * size * (t + 1) * tdim or t == n_t - 1, (bsize, size, t, y)
