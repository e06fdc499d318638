"""Fair A/B of two tuning settings: the two are timed alternately (A B A B ...), several rounds, so that clock / thermal
drift and first-measurement effects cancel.   MSDA_AB=1 python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=15 [--dtype f32]"""
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
ap.add_argument("--mode", default="model")
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--what", default="fwd", choices=["fwd", "bwd"])
ap.add_argument("--head-dim", type=int, default=32)
ap.add_argument("--heads", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda:0")
wl = W.config(1, loc_mode=args.mode, head_dim=args.head_dim, heads=args.heads, dtype={"f32": torch.float32, "bf16": torch.bfloat16}[args.dtype])
d = W.make_inputs(wl, device=dev)
a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
fn = (lambda: torch.ops.msda.forward(*a5, 64)) if args.what == "fwd" else (lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64))


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
print(json.dumps({"what": args.what, "dtype": args.dtype, "mode": args.mode, "head_dim": args.head_dim, "heads": args.heads,
                  **{k: {"median_ms": round(statistics.median(v), 4), "min_ms": round(min(v), 4), "all": [round(x, 4) for x in v]}
                     for k, v in res.items()}}))
