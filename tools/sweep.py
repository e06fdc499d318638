"""Tuning sweep on the GPU box: time forward / backward under different tuning knobs.
    python tools/sweep.py --set fwd_variant=11,fwd_pipe=2 --set fwd_variant=1 [--dtypes f32,bf16] [--modes model]
Each --set is one configuration (comma-separated key=value for msda_set_tuning); prints one JSON
line per (loc_mode, dtype, configuration)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402

KEYS = ("fwd_variant", "bwd_variant", "fwd_pipe", "bwd_pipe")


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--set", action="append", default=[])
    ap.add_argument("--dtypes", default="f32")
    ap.add_argument("--modes", default="model")
    ap.add_argument("--cfg", type=int, default=1)
    ap.add_argument("--queries", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = msda._lib
    sets = a.set or [""]
    for mode in a.modes.split(","):
        for dt in a.dtypes.split(","):
            dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
            over = dict(batch=a.batch, loc_mode=mode, dtype=dtype)
            if a.queries:
                over["num_queries"] = a.queries
            wl = W.config(a.cfg, **over)
            d = W.make_inputs(wl, device=dev)
            ab = W.algorithmic_bytes(wl)
            a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
            for cfg in sets:
                for k in KEYS:
                    L.set_tuning(k, -1)
                for kv in filter(None, cfg.split(",")):
                    k, v = kv.split("=")
                    L.set_tuning(k, int(v))
                f = timeit(lambda: torch.ops.msda.forward(*a5, 64), a.iters)
                b = timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), a.iters)
                print(json.dumps(dict(wl=wl.name, mode=mode, dtype=dt, tuning=cfg or "default", fwd_ms=round(f, 4), bwd_ms=round(b, 4),
                                      fwd_GBps=round(ab["fwd"] / f / 1e6, 1), bwd_GBps=round(ab["bwd"] / b / 1e6, 1),
                                      total_GBps=round(ab["total"] / (f + b) / 1e6, 1))), flush=True)
            del d, a5
            torch.cuda.empty_cache()
    for k in KEYS:
        L.set_tuning(k, -1)


if __name__ == "__main__":
    main()
