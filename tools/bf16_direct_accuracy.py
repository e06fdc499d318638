"""A/B: bf16 backward at the encoder shape with fp32 accumulation (default) vs straight REDG.E.ADD.BF16x4 into the bf16
gradient (msda_set_tuning("bf16_direct", 1000)): time and rel-L2 of grad_value against the fp64 oracle, per level.
    python tools/bf16_direct_accuracy.py > gpurun_out/r02_bf16_direct_accuracy.jsonl"""
import sys, torch, json
sys.path.insert(0, "/root/repo")
import monosowa_b200 as msda
from monosowa_b200 import workloads as W
from oracle import msda_oracle as O
from tools.sweep import timeit
dev = torch.device("cuda:0")
for thr in (-1, 1000):
    msda._lib.set_tuning("bf16_direct", thr)
    for mode in ("model", "init"):
        wl = W.config(1, dtype=torch.bfloat16, loc_mode=mode)
        d = W.make_inputs(wl, device=dev)
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
        gv, gl, ga = torch.ops.msda.backward(*a5, d["grad_out"], 64)
        t = timeit(lambda: torch.ops.msda.backward(*a5, d["grad_out"], 64), 10)
        i = 3
        sl = lambda x: x[i:i+1].detach().cpu().double()
        rgv, rgl, rga = O.backward_c(sl(d["value"]), d["shapes"].cpu(), d["lsi"].cpu(), sl(d["loc"]), sl(d["attn"]), sl(d["grad_out"]))
        lv = [0, 7680, 9600, 10080, 10200]
        per = [round(O.rel_l2(gv[i:i+1, lv[k]:lv[k+1]], rgv[:, lv[k]:lv[k+1]]), 5) for k in range(4)]
        print(json.dumps(dict(bf16_direct=thr, mode=mode, bwd_ms=round(t, 4), gv_rel_l2=O.rel_l2(gv[i:i+1], rgv), per_level=per)))
