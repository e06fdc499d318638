#!/bin/bash
# forward accumulation as four chained FMAs per channel (fwd_pipe=40) against acc += (sum of four products)
mkdir -p gpurun_out
export MSDA_AB=1
python - <<'PY'
import sys, torch
sys.path.insert(0, "/root/repo")
import monosowa_b200 as msda
from monosowa_b200 import workloads as W
from oracle import msda_oracle as O
dev = torch.device("cuda:0")
wl = W.config(1, batch=2)
d = W.make_inputs(wl, device=dev)
a5 = (d["value"], d["shapes"], d["lsi"], d["loc"], d["attn"])
ref = O.forward_c(d["value"].double().cpu(), d["shapes"].cpu(), d["lsi"].cpu(), d["loc"].double().cpu(), d["attn"].double().cpu())
for fp in (-1, 40):
    msda._lib.set_tuning("fwd_pipe", fp)
    out = torch.ops.msda.forward(*a5, 64)
    print("fwd_pipe", fp, "rel-L2 vs fp64 oracle", O.rel_l2(out, ref))
PY
O=gpurun_out/r02_fwd_chained_fma_interleaved.jsonl; : > $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=40 --rounds 8 >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=40 --rounds 8 --mode uniform >> $O
cat $O
