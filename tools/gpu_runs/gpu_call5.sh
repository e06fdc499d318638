#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_step_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
tail -25 gpurun_out/r02_pytest5.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
