#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_bench_gpu.py tests/test_msda_gpu.py -m gpu -q -k "contract or lane_widths or fuzz or variant" 2>&1 | tail -4 | cut -c1-400
O=gpurun_out/r02_bwd_d64_interleaved.jsonl; : > $O
python tools/ab_interleaved.py bwd_variant=-1 bwd_variant=11 --what bwd --head-dim 64 --heads 4 >> $O
python tools/ab_interleaved.py bwd_variant=-1 bwd_variant=11 --what bwd --head-dim 64 --heads 4 --dtype bf16 >> $O
python tools/ab_interleaved.py bwd_variant=11,bwd_pipe=-1 bwd_variant=11,bwd_pipe=4 --what bwd --head-dim 64 --heads 4 --dtype bf16 >> $O
cat $O
