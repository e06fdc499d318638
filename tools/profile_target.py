"""Small ncu target: a few forward+backward launches of the headline workload (configs[1]).
    python tools/profile_target.py [--dtype f32|bf16] [--iters 3] [--mode model|uniform]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402,F401
from monosowa_b200 import workloads as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f32")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--mode", default="model")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--set", default="", help="comma-separated key=value for msda_set_tuning")
a = ap.parse_args()
dev = torch.device("cuda:0")
for kv in filter(None, a.set.split(",")):
    k_, v_ = kv.split("=")
    msda._lib.set_tuning(k_, int(v_))
wl = W.config(1, batch=a.batch, loc_mode=a.mode, dtype={"f32": torch.float32, "bf16": torch.bfloat16}[a.dtype])
d = W.make_inputs(wl, device=dev)
a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
for _ in range(a.iters):
    out = torch.ops.msda.forward(*a5, 64)
    g = torch.ops.msda.backward(*a5, d["grad_out"], 64)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()), float(g[0].float().abs().mean()))
