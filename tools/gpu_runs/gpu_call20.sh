#!/bin/bash
# fused flavours at the encoder shape: unfused vs fused timing, and the fused binned backward at 2 CTAs/SM (128 registers)
mkdir -p gpurun_out
export MSDA_AB=1
O=gpurun_out/r02_fused_ab.jsonl; : > $O
python tools/fused_ab.py bwd_pipe=-1 bwd_pipe=92 --what bwd >> $O
python tools/fused_ab.py bwd_pipe=-1 bwd_pipe=92 --what bwd --dtype bf16 >> $O
python tools/fused_ab.py fwd_pipe=-1 fwd_pipe=4 --what fwd >> $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=92 --what bwd >> $O
cat $O
