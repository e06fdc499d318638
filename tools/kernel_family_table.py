import sys, os, json, torch
sys.path.insert(0, "/root/repo")
import monosowa_b200 as msda
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
