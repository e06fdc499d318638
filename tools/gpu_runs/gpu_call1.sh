#!/bin/bash
# first GPU pass of round 2: correctness of the tile kernels, then A/B timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt
timeout 1500 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 300 python tools/sweep.py --set "" --set fwd_variant=11,bwd_variant=11 --dtypes f32,bf16 --modes model,uniform,init > gpurun_out/r02_sweep_tile_vs_rec.jsonl 2>&1
cat gpurun_out/r02_sweep_tile_vs_rec.jsonl
MSDA_AB=1 timeout 300 python tools/sweep.py --set fwd_pipe=0 --set fwd_pipe=1 --set fwd_pipe=2 --set fwd_pipe=3 --set fwd_pipe=4 --set fwd_pipe=5 --set fwd_pipe=6 --set fwd_pipe=7 --set fwd_pipe=8 --dtypes f32,bf16 > gpurun_out/r02_sweep_fwd_tile_flavours.jsonl 2>&1
cat gpurun_out/r02_sweep_fwd_tile_flavours.jsonl
