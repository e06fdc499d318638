#!/bin/bash
# record layout without write conflicts for 4-lane groups: forward timings (fwd_pipe=4 is the unchanged 8-lane flavour)
mkdir -p gpurun_out
export MSDA_AB=1
O=gpurun_out/r02_fwd_layout_interleaved.jsonl; : > $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 --mode uniform >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=4 --dtype bf16 >> $O
python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=27 >> $O
cat $O
unset MSDA_AB
python -m pytest tests/test_msda_gpu.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
