"""Measure every BASELINE.json MSDA config (and the config-4 shape sweep) on one GPU.
    python tools/config_table.py > gpurun_out/config_table.jsonl
One JSON line per workload: fwd/bwd ms, algorithmic GB/s, kernels used."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402
from tools.sweep import timeit  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    wls = [W.config(1), W.config(1, loc_mode="uniform"), W.config(1, dtype=torch.bfloat16),
           W.config(2, num_queries=50), W.config(2, num_queries=550),
           W.config(2, num_queries=50, dtype=torch.float32), W.config(2, num_queries=550, dtype=torch.float32)]
    wls += W.sweep_config5(batch=4)
    lib = msda._lib.lib
    for wl in wls:
        d = W.make_inputs(wl, device=dev)
        ab = W.algorithmic_bytes(wl)
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
        f = timeit(lambda: torch.ops.msda.forward(*a5, 64), 20)
        b = timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), 20)
        bf = wl.dtype == torch.bfloat16
        print(json.dumps(dict(
            workload=wl.name, dtype=str(wl.dtype).replace("torch.", ""), batch=wl.batch, S=wl.S, Lq=wl.Lq, loc_mode=wl.loc_mode,
            fwd_us=round(f * 1e3, 1), bwd_us=round(b * 1e3, 1), fwd_GBps=round(ab["fwd"] / f / 1e6, 1),
            bwd_GBps=round(ab["bwd"] / b / 1e6, 1), total_GBps=round(ab["total"] / (f + b) / 1e6, 1),
            frac_of_measured_hbm=round(ab["total"] / (f + b) / 1e6 / 6559.4, 4),
            alg_MB=round(ab["total"] / 1e6, 1),
            kernels=[msda._lib.describe("forward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq),
                     msda._lib.describe("backward", wl.dtype, wl.batch, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq)])), flush=True)
        del d, a5
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
