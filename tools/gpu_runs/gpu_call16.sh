#!/bin/bash
# e2e at one GPU: pipeline depth and chunk size (host step with one-image first / last chunks, forward started after its
# three inputs, `out` copied out behind the forward)
mkdir -p gpurun_out
python -m pytest tests/test_msda_gpu.py -m gpu -q -k "host_buffer" 2>&1 | tail -2 | cut -c1-300
O=gpurun_out/r02_n1_e2e_stages.txt; : > $O
for f in "--e2e-stages 3 --e2e-images-per-chunk 1" "--e2e-stages 3 --e2e-images-per-chunk 2" "--e2e-stages 8 --e2e-images-per-chunk 2" "--e2e-stages 4 --e2e-images-per-chunk 3" "--e2e-stages 4 --e2e-images-per-chunk 4"; do
  python bench.py --steps 20 --warmup 5 --sustain-steps 20 --no-train-step --no-other-configs --no-cpu-baseline --no-ref-cuda $f 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('$f', 'e2e ms', round(e['ms_per_step'],2), 'GB/s', round(e['value'],1), 'duplex floor ms', round(e['pcie']['duplex_ms'],2), 'frac', round(e['frac_of_pcie_floor'],3), 'autograd-api ms', round(e['autograd_api_ms_per_step'],2))" >> $O
done
cat $O
