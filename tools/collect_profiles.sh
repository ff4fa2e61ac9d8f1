#!/bin/bash
# Runs on the GPU box (under gpurun): plain run first, then the ncu passes of B200_PROFILING.md.
# Outputs land in gpurun_out/; summarise them into profiles/ with tools/summarize_launches.py.
set -u
R=${1:-r01}
CMD="python bench.py --steps 1 --warmup 1"
export B200CLIP_GRAPH=0
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
# every launch of the timed step with its device time (warm-up step = first 511 launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 511 -c 511 --csv \
    --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# DRAM traffic of every GEMM launch of that step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm_bf16_kernel -s 290 -c 300 --csv --log-file gpurun_out/${R}_gemm_dram.csv $CMD > gpurun_out/${R}_ncu_dram.log 2>&1
echo "gemm dram rc=$?"
# full capture of three GEMM launches (forward c_fc, mid-forward)
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 330 -c 3 \
    -o gpurun_out/${R}_gemm_full $CMD > gpurun_out/${R}_ncu_full.log 2>&1
echo "full rc=$?"
