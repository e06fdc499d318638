#!/bin/bash
# final refresh of the round-2 numbers after the 8-channel forward: bench lines (own arm, reference arm), config table,
# module bench, launch list, ncu --set full of the two shipped kernels, smoke
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "reference arm rc=$?"
python tools/config_table.py > gpurun_out/r02_config_table.jsonl 2> gpurun_out/r02_config_table.err; echo "config table rc=$?"
python tools/module_bench.py > gpurun_out/r02_module_bench.jsonl 2> gpurun_out/r02_module_bench.err; echo "module bench rc=$?"
bash tools/gpu_runs/gpu_call6.sh
cut -c1-1200 gpurun_out/r02_bench.json
