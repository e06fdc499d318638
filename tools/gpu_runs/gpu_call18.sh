#!/bin/bash
# binned backward with 8 channels per lane at D = 64: parity, then interleaved timings against the 4-channel flavour
mkdir -p gpurun_out
python -m pytest tests/test_msda_gpu.py -m gpu -q -x -k "long_query or fused or fuzz or lane_widths or variant" 2>&1 | tail -3 | cut -c1-300
python tools/fuzz_gpu.py --cases 200 --seed 3 2>&1 | tail -2
export MSDA_AB=1
O=gpurun_out/r02_bwd_bin_d64_interleaved_b.jsonl; : > $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=4 --what bwd --head-dim 64 --heads 4 >> $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=83 --what bwd --head-dim 64 --heads 4 >> $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=83 --what bwd --head-dim 64 --heads 4 --dtype bf16 >> $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=4 --what bwd --head-dim 64 --heads 4 --dtype bf16 >> $O
python tools/ab_interleaved.py bwd_pipe=-1 bwd_pipe=4 --what bwd --head-dim 64 --heads 8 >> $O
cat $O
