#!/bin/bash
# where do the forward's extra DRAM reads come from (570 MB vs 418 MB compulsory)?  DRAM bytes and time per flavour:
# -1 shipped (8 channels per lane, evict-normal input streams), 4 four channels per lane (evict-first streams),
# 24 allocating gathers, 30 evict_last gathers, 31 evict-first streams, 33 evict-first streams with an L2::128B hint
mkdir -p gpurun_out
export MSDA_AB=1
O=gpurun_out/r02_fwd_dram_by_flavour.txt; : > $O
for s in ${FLAVOURS:-"fwd_pipe=-1" "fwd_pipe=4" "fwd_pipe=24" "fwd_pipe=30" "fwd_pipe=31" "fwd_pipe=33"}; do
  echo "== $s" >> $O
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fwd_rec -c 1 python tools/profile_target.py --iters 1 --set "$s" 2>&1 | grep -E "fwd_rec_kernel<|duration|dram__|lts__" >> $O
done
cat $O
T=gpurun_out/r02_fwd_stream_policy_interleaved.jsonl; : > $T
for p in 31; do python tools/ab_interleaved.py fwd_pipe=-1 fwd_pipe=$p >> $T; done
cat $T
bash tools/gpu_runs/gpu_call6.sh
