#!/bin/bash
# ncu evidence for the bench step (run under gpurun, 1 GPU).  Outputs under gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --distinct-clips 1"
# every launch of one timed step + warm-up with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
for k in gemm_tcgen05 clip_preprocess attention_kernel layernorm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
