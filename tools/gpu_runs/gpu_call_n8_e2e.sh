#!/bin/bash
for c in 1 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$c bench.py --gpus 8 --steps 5 --warmup 3 --no-train-step --no-other-configs --no-cpu-baseline --no-ref-cuda --sustain-steps 1 --e2e-images-per-chunk $c > gpurun_out/r02_n8_e2e_chunk$c.json 2>/dev/null
python - <<PY
import json
l=json.loads(open("gpurun_out/r02_n8_e2e_chunk$c.json").read().strip().splitlines()[-1])
print("chunk", $c, "e2e ms", round(l["e2e"]["ms_per_step"],2), "GB/s", round(l["e2e"]["value"],1), "floor ms", round(l["e2e"]["pcie"]["duplex_ms"],2), "h2d", round(l["e2e"]["pcie"]["h2d_GBps"],1), "d2h", round(l["e2e"]["pcie"]["d2h_GBps"],1), "autograd-api ms", round(l["e2e"]["autograd_api_ms_per_step"],2))
PY
done
