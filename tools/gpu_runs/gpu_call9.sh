#!/bin/bash
# record backward with 8 channels per lane (measurement build): parity, then interleaved timings at configs[1]
mkdir -p gpurun_out
python -m pytest tests/test_bench_gpu.py -m gpu -q -k contract 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-400
export MSDA_AB=1
python - <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
import monosowa_b200 as msda
from monosowa_b200 import workloads as W
dev = torch.device("cuda:0")
for dt in (torch.float32, torch.bfloat16):
    for hd, heads in ((32, 8), (64, 4)):
        wl = W.config(1, batch=2, dtype=dt, head_dim=hd, heads=heads)
        d = W.make_inputs(wl, device=dev)
        a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
        msda._lib.set_tuning("bwd_variant", 11)
        ref = torch.ops.msda.backward(*a5, d["grad_out"], 64)
        for bp in (82, 83):
            msda._lib.set_tuning("bwd_pipe", bp)
            out = torch.ops.msda.backward(*a5, d["grad_out"], 64)
            print(dt, hd, bp, "rel diff vs 4-channel record kernel: gv %.2e gl %.2e ga %.2e" % tuple(
                ((o.float() - r.float()).norm() / r.float().norm()).item() for o, r in zip(out, ref)))
        msda._lib.set_tuning("bwd_pipe", -1); msda._lib.set_tuning("bwd_variant", -1)
PY
O=gpurun_out/r02_bwd_rec_cpl8_interleaved.jsonl; : > $O
for p in 82 83; do python tools/ab_interleaved.py bwd_variant=11,bwd_pipe=-1 bwd_variant=11,bwd_pipe=$p --what bwd >> $O; done
python tools/ab_interleaved.py bwd_variant=11,bwd_pipe=-1 bwd_variant=11,bwd_pipe=83 --what bwd --dtype bf16 >> $O
python tools/ab_interleaved.py bwd_variant=11,bwd_pipe=-1 bwd_variant=11,bwd_pipe=83 --what bwd --head-dim 64 --heads 4 >> $O
cat $O
