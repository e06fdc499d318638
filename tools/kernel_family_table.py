"""Shipped default vs record kernels vs tile kernels (measurement build: MSDA_AB=1) on configs[1] and the configs[4]
shapes:   MSDA_AB=1 python tools/kernel_family_table.py > gpurun_out/kernel_family_table.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W
from tools.sweep import timeit
dev = torch.device("cuda:0")
wls = [W.config(1), W.config(1, dtype=torch.bfloat16)] + W.sweep_config5(batch=4)
for wl in wls:
    d = W.make_inputs(wl, device=dev)
    a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
    row = dict(wl=wl.name)
    for name, fv, bv in (("default", -1, -1), ("rec", 11, 11), ("tile", 12, 20)):
        msda._lib.set_tuning("fwd_variant", fv); msda._lib.set_tuning("bwd_variant", bv)
        row[name + "_fwd_us"] = round(timeit(lambda: torch.ops.msda.forward(*a5, 64), 10) * 1e3, 1)
        row[name + "_bwd_us"] = round(timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), 10) * 1e3, 1)
    msda._lib.set_tuning("fwd_variant", -1); msda._lib.set_tuning("bwd_variant", -1)
    print(json.dumps(row), flush=True)
    del d, a5; torch.cuda.empty_cache()
