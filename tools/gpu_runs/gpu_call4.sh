#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -4 gpurun_out/r02_pytest4.log
MSDA_AB=1 timeout 1500 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short --maxfail=12 -p no:cacheprovider -k "tile or variant or long_query or fused or guard or nan" > gpurun_out/r02_pytest4_ab.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest4_ab.log
tail -4 gpurun_out/r02_pytest4_ab.log
python __graft_entry__.py smoke 2>&1 | tail -4
