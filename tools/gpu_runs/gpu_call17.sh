#!/bin/bash
# final bench lines of round 2 (own arm and reference arm), as the driver runs them
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "reference arm rc=$?"
cut -c1-300 gpurun_out/r02_bench.json
