#!/bin/bash
# ncu evidence for the bench step (run under gpurun, 1 GPU).  Outputs under gpurun_out/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-shots --no-secondary --no-gpu-library --distinct-clips 4"
# every launch of warm-up + one timed step with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_2cta -s 8 -c 3 -f -o gpurun_out/prof_gemm_tcgen05_2cta $B > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attention_tc2 -s 4 -c 1 -f -o gpurun_out/prof_attention_tc2 $B > gpurun_out/ncu_attention.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 8 -c 2 -f -o gpurun_out/prof_layernorm_kernel $B > gpurun_out/ncu_layernorm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:clip_preprocess_tc -s 1 -c 1 -f -o gpurun_out/prof_clip_preprocess_tc $B > gpurun_out/ncu_clip_preprocess.log 2>&1
ls -la gpurun_out/*.ncu-rep
