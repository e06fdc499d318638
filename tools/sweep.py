"""Tuning sweep on the GPU box: time forward / backward for every launch variant.
    python tools/sweep.py [--batch 16] [--iters 10]
Prints one JSON line per (loc_mode, dtype, order, threads)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402


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
    ap.add_argument("--orders", default="0,1")
    ap.add_argument("--threads", default="128,256,512")
    ap.add_argument("--dtypes", default="f32,bf16")
    ap.add_argument("--modes", default="model,uniform")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    L = msda._lib
    for mode in a.modes.split(","):
        for dt in a.dtypes.split(","):
            dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dt]
            wl = W.config(1, batch=a.batch, loc_mode=mode, dtype=dtype)
            d = W.make_inputs(wl, device=dev)
            ab = W.algorithmic_bytes(wl)
            a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
            for order in [int(x) for x in a.orders.split(",")]:
                for th in [int(x) for x in a.threads.split(",")]:
                    L.set_tuning("fwd_variant", order); L.set_tuning("bwd_variant", order); L.set_tuning("block_threads", th)
                    f = timeit(lambda: torch.ops.msda.forward(*a5, 64), a.iters)
                    b = timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), a.iters)
                    print(json.dumps(dict(mode=mode, dtype=dt, order=order, threads=th, fwd_ms=round(f, 4), bwd_ms=round(b, 4),
                                          fwd_GBps=round(ab["fwd"] / f / 1e6, 1), bwd_GBps=round(ab["bwd"] / b / 1e6, 1),
                                          total_GBps=round(ab["total"] / (f + b) / 1e6, 1))), flush=True)
            del d, a5
            torch.cuda.empty_cache()
    for k in ("fwd_variant", "bwd_variant", "block_threads"):
        L.set_tuning(k, -1)


if __name__ == "__main__":
    main()
