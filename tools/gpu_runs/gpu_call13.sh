#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_msda_gpu.py -m gpu -q -k "fused or lane_widths or long_query" 2>&1 | tail -3 | cut -c1-300
bash tools/gpu_runs/gpu_call6.sh
