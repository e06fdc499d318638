"""Where does the binned backward start to pay?  Record (bwd_variant=11) vs binned (21) backward at the KITTI encoder
shape over the batch size, i.e. over the number of (image, head, 256-query chunk) work items:
    python tools/binned_crossover.py > gpurun_out/binned_crossover.jsonl"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monosowa_b200 as msda  # noqa: E402
from monosowa_b200 import workloads as W  # noqa: E402
from tools.sweep import timeit  # noqa: E402

dev = torch.device("cuda:0")
for shapes, tag in ((W.KITTI, "kitti"), (W.WAYMO, "waymo")):
    for n in ((2, 4, 6, 8, 10, 12, 16) if tag == "kitti" else (1, 2, 4)):
        wl = W.Workload(name=f"{tag}_b{n}", shapes=shapes, batch=n, queries="encoder", seed=1235)
        d = W.make_inputs(wl, device=dev)
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
        row = dict(workload=wl.name, work_items=n * wl.heads * ((wl.Lq + 255) // 256))
        for name, v in (("rec", 11), ("binned", 21)):
            msda._lib.set_tuning("bwd_variant", v)
            row[name + "_us"] = round(timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), 20) * 1e3, 1)
        msda._lib.set_tuning("bwd_variant", -1)
        row["default"] = msda._lib.describe("backward", wl.dtype, n, wl.heads, wl.head_dim, wl.L, wl.points, wl.Lq)
        print(json.dumps(row), flush=True)
        del d, a5
        torch.cuda.empty_cache()
