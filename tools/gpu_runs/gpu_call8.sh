#!/bin/bash
# after the 8-channel forward: full GPU suite (product build, then the measurement-build subset)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8
