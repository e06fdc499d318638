#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider -k "tile or variant or nan or fused or config2 or guard or host_buffer or config1" > gpurun_out/r02_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -4 gpurun_out/r02_pytest3.log
timeout 300 python tools/sweep.py --set "" --set fwd_variant=11,bwd_variant=11 --dtypes f32,bf16 --modes model,uniform,init > gpurun_out/r02_sweep_tile_vs_rec_v3.jsonl 2>&1
cat gpurun_out/r02_sweep_tile_vs_rec_v3.jsonl
ncu --set full --clock-control none --import-source on -k regex:bwd_tile -c 1 -f -o gpurun_out/r02_bwd_tile_v3 python tools/profile_target.py --iters 1 > gpurun_out/r02_ncu3.log 2>&1; tail -2 gpurun_out/r02_ncu3.log
