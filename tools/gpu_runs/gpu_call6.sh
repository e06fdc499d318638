#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-train-step --no-other-configs --no-cpu-baseline --no-ref-cuda --sustain-steps 3 > gpurun_out/r02_bench_for_launches.json 2>/dev/null
echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-train-step --no-other-configs --no-cpu-baseline --no-ref-cuda --sustain-steps 3 > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python tools/profile_target.py --iters 1 > /dev/null && ncu --set full --clock-control none --import-source on -k regex:"fwd_rec|bwd_bin" -c 2 -f -o gpurun_out/r02_final python tools/profile_target.py --iters 1 > gpurun_out/r02_ncu_final.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r02_ncu_final.log
