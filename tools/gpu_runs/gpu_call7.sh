#!/bin/bash
# 8 channels per lane in the forward: bitwise check against the 4-channel flavour, then interleaved timings
mkdir -p gpurun_out
export MSDA_AB=1
python tools/_chk8.py 2>&1 | tail -40
O=gpurun_out/r02_fwd_cpl8_interleaved.jsonl; : > $O
for m in model uniform init; do python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 --mode $m >> $O; done
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=27 >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=24 >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 --head-dim 64 --heads 4 >> $O
for p in 4 22 25 26; do python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=$p --dtype bf16 >> $O; done
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 --dtype bf16 --head-dim 64 --heads 4 >> $O
cat $O
