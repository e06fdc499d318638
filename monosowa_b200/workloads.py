"""Synthetic MSDA workloads (BASELINE.json configs) and the algorithmic-byte model.

Pure shape/tensor plumbing shared by tests and bench.py; no kernels, no oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


def pyramid(img_h: int, img_w: int, levels: int = 4) -> List[Tuple[int, int]]:
    """Feature-map sizes MonoDETR's backbone produces for an image: strides 8/16/32 from ResNet
    (ceil division per stride-2 stage) and one more 3x3-s2-p1 conv per extra level
    (reference monodetr.py:161-176).  384x1280 -> (48,160),(24,80),(12,40),(6,20)."""
    h, w = img_h, img_w
    for _ in range(3):
        h, w = math.ceil(h / 2), math.ceil(w / 2)
    out = [(h, w)]
    for _ in range(levels - 1):
        h, w = math.ceil(h / 2), math.ceil(w / 2)
        out.append((h, w))
    return out


@dataclass
class Workload:
    name: str
    shapes: List[Tuple[int, int]]
    batch: int
    queries: str                 # "encoder" (Lq = S, queries are the pixel grid) or "decoder"
    num_queries: int = 0         # decoder only
    heads: int = 8
    head_dim: int = 32
    points: int = 4
    dtype: torch.dtype = torch.float32
    loc_mode: str = "model"      # "model" (pixel-centre refs + N(0,2px) offsets), "uniform", or "init"
                                 # ("init": the module's initial offset pattern, identical for all queries
                                 #  -- reference ms_deform_attn.py:106-115; "init+noise": plus N(0, 0.5 px))
    seed: int = 1234
    note: str = ""
    S: int = field(init=False)
    Lq: int = field(init=False)

    def __post_init__(self):
        self.S = sum(h * w for h, w in self.shapes)
        self.Lq = self.S if self.queries == "encoder" else self.num_queries

    @property
    def L(self):
        return len(self.shapes)


KITTI = pyramid(384, 1280)            # [(48,160),(24,80),(12,40),(6,20)], S = 10200
KITTI360 = pyramid(376, 1408)         # (47,176)..., S = 11044
WAYMO = pyramid(1280, 1920)           # (160,240)..., S = 51000
ALT640 = pyramid(640, 960)            # S = 12750


def config(index: int, **over) -> Workload:
    """BASELINE.json `configs[index]` (0-based)."""
    table = {
        0: dict(name="cfg0_kitti_enc_b2_f32_cpu", shapes=KITTI, batch=2, queries="encoder"),
        1: dict(name="cfg1_kitti_enc_b16_f32", shapes=KITTI, batch=16, queries="encoder"),
        2: dict(name="cfg2_kitti_dec_b16_bf16", shapes=KITTI, batch=16, queries="decoder", num_queries=50,
                dtype=torch.bfloat16),
    }
    kw = dict(table[index])
    kw["seed"] = 1234 + index
    kw.update(over)
    return Workload(**kw)


def sweep_config5(batch: int = 4):
    """BASELINE.json configs[4]: large-image shape sweep (one replica per GPU)."""
    out = []
    for tag, shp in (("kitti360_376x1408", KITTI360), ("waymo_1280x1920", WAYMO), ("alt_640x960", ALT640)):
        for dt, dtag in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
            out.append(Workload(name=f"cfg4_{tag}_b{batch}_{dtag}", shapes=shp, batch=batch, queries="encoder",
                                dtype=dt, seed=1238))
    return out


def level_tensors(shapes, device):
    sh = torch.as_tensor(shapes, dtype=torch.long, device=device)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return sh, lsi


def encoder_reference_points(shapes, device, dtype=torch.float32):
    """Pixel centres of every level, normalised, replicated to all levels -- what
    VisualEncoder.get_reference_points yields with valid_ratio 1
    (reference depthaware_transformer.py:363-376).  Returns (S, L, 2) in (x, y)."""
    pts = []
    for h, w in shapes:
        ys = (torch.arange(h, device=device, dtype=dtype) + 0.5) / h
        xs = (torch.arange(w, device=device, dtype=dtype) + 0.5) / w
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack((gx.reshape(-1), gy.reshape(-1)), -1))
    ref = torch.cat(pts, 0)
    return ref[:, None, :].expand(-1, len(shapes), -1)


def make_inputs(wl: Workload, device="cpu", seed=None, requires_grad=False):
    """Seeded inputs: value ~ N(0,1), attention = softmax(N(0,1)) over L*P, grad_out ~ N(0,1),
    locations per `wl.loc_mode`.  loc/attn are fp32 (fp64 if wl.dtype is fp64)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(wl.seed if seed is None else seed)
    ct = torch.float64 if wl.dtype == torch.float64 else torch.float32
    N, S, M, D, L, P, Lq = wl.batch, wl.S, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq
    sh, lsi = level_tensors(wl.shapes, dev)
    value = torch.randn(N, S, M, D, generator=g, device=dev, dtype=ct).to(wl.dtype)
    wh = torch.stack([sh[:, 1], sh[:, 0]], -1).to(ct)                     # (L,2) as (W,H)
    if wl.loc_mode == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g, device=dev, dtype=ct)
    else:
        if wl.queries == "encoder":
            ref = encoder_reference_points(wl.shapes, dev, ct)[None].expand(N, -1, -1, -1)
        else:
            ref = torch.rand(N, Lq, 1, 2, generator=g, device=dev, dtype=ct).mul_(0.9).add_(0.05).expand(-1, -1, L, -1)
        if wl.loc_mode.startswith("init"):
            theta = torch.arange(M, device=dev, dtype=ct) * (2.0 * math.pi / M)
            ring = torch.stack([theta.cos(), theta.sin()], -1)
            ring = ring / ring.abs().max(-1, keepdim=True)[0]                                   # (M, 2)
            steps = torch.arange(1, P + 1, device=dev, dtype=ct)
            off_px = (ring[:, None, None, :] * steps[None, None, :, None]).expand(M, L, P, 2)
            off_px = off_px[None, None].expand(N, Lq, M, L, P, 2)
            if wl.loc_mode == "init+noise":
                off_px = off_px + torch.randn(N, Lq, M, L, P, 2, generator=g, device=dev, dtype=ct).mul_(0.5)
        else:
            off_px = torch.randn(N, Lq, M, L, P, 2, generator=g, device=dev, dtype=ct).mul_(2.0)
        loc = ref[:, :, None, :, None, :] + off_px / wh[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g, device=dev, dtype=ct), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g, device=dev, dtype=ct).to(wl.dtype)
    loc, attn = loc.contiguous(), attn.contiguous()
    if requires_grad:
        value.requires_grad_(True); loc.requires_grad_(True); attn.requires_grad_(True)
    return dict(value=value, shapes=sh, lsi=lsi, loc=loc, attn=attn, grad_out=grad_out)


def live_corner_rows(d, wl: Workload) -> int:
    """Number of (sample, corner) rows that carry weight: samples inside the (-1,H)x(-1,W) window (reference
    ms_deform_im2col_cuda.cuh:288) times their corners inside the map -- the rows a forward has to gather."""
    loc = d["loc"].float()
    total = 0
    for l, (h, w) in enumerate(wl.shapes):
        px = loc[:, :, :, l, :, 0] * w - 0.5
        py = loc[:, :, :, l, :, 1] * h - 0.5
        inside = (px > -1) & (py > -1) & (px < w) & (py < h)
        x0, y0 = px.floor(), py.floor()
        nx = ((x0 >= 0) & inside).long() + ((x0 + 1 <= w - 1) & inside).long()
        ny = ((y0 >= 0) & inside).long() + ((y0 + 1 <= h - 1) & inside).long()
        total += int((nx * ny).sum().item())
    return total


def algorithmic_bytes(wl: Workload):
    """Compulsory traffic of one forward / one backward: every input read once,
    every output written once; a sparsely gathered tensor counts min(dense, gathered)."""
    ev = torch.empty((), dtype=wl.dtype).element_size()
    el = 8 if wl.dtype == torch.float64 else 4
    N, S, M, D, L, P, Lq = wl.batch, wl.S, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq
    V = N * S * M * D * ev
    G = N * Lq * M * L * P * 4 * D * ev
    Vt = min(V, G)
    Bloc = N * Lq * M * L * P * 2 * el
    Baw = N * Lq * M * L * P * el
    O = N * Lq * M * D * ev
    fwd = Vt + Bloc + Baw + O
    bwd = Vt + Bloc + Baw + O + V + Bloc + Baw
    return dict(fwd=fwd, bwd=bwd, total=fwd + bwd, gather=G)
