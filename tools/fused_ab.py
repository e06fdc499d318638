"""Interleaved A/B of two tuning settings on the FUSED ops (raw offsets + logits + reference points) at the encoder shape of
BASELINE.json configs[1]:   MSDA_AB=1 python tools/fused_ab.py bwd_pipe=-1 bwd_pipe=92 [--what bwd] [--dtype f32]"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402
from tools.sweep import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("a")
ap.add_argument("b")
ap.add_argument("--dtype", default="f32")
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--what", default="bwd", choices=["fwd", "bwd"])
args = ap.parse_args()
dev = torch.device("cuda:0")
dt = {"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype]
wl = W.config(1, dtype=dt)
d = W.make_inputs(wl, device=dev)
N, Lq, M, L, P = wl.batch, wl.Lq, wl.heads, wl.L, wl.points
g = torch.Generator(device="cpu").manual_seed(5)
ref = W.encoder_reference_points(wl.shapes, dev)[None].expand(N, -1, -1, -1).contiguous()
offsets = (torch.randn(N, Lq, M, L, P, 2, generator=g) * 2.0).to(dev)             # pixels, as the module's Linear emits them
logits = torch.randn(N, Lq, M, L * P, generator=g).to(dev)
a6 = (d["value"], d["shapes"], d["lsi"], ref, offsets, logits)
fn = (lambda: torch.ops.msda.forward_fused(*a6)) if args.what == "fwd" else \
     (lambda: torch.ops.msda.backward_fused(*a6, d["grad_out"], False))


def apply(cfg):
    for kv in cfg.split(","):
        k, v = kv.split("=")
        msda._lib.set_tuning(k, int(v))


res = {args.a: [], args.b: []}
for _ in range(3):
    fn()
for r in range(args.rounds):
    for cfg in ((args.a, args.b) if r % 2 == 0 else (args.b, args.a)):
        apply(cfg)
        res[cfg].append(timeit(fn, 30))
print(json.dumps({"op": "fused " + args.what, "dtype": args.dtype,
                  **{k: {"median_ms": round(statistics.median(v), 4), "min_ms": round(min(v), 4), "all": [round(x, 4) for x in v]}
                     for k, v in res.items()}}))
