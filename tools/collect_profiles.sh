#!/bin/bash
# Runs on the GPU box (under gpurun): the plain run first, then ONE ncu pass of B200_PROFILING.md over
# a single profiled training step (tools/profile_one_step.py).  usage: collect_profiles.sh <pass> [tag]
#   launches : gpu__time_duration.sum of every launch of the step
#   dram     : dram__bytes_read/write.sum + duration of every GEMM launch of the step
#   full     : --set full --import-source on for three forward GEMM launches
# Outputs land in gpurun_out/; summarise them into profiles/ with tools/summarize_launches.py.
set -u
PASS=${1:-launches}
R=${2:-r01}
CMD="python tools/profile_one_step.py ${PAIRS:-1024}"
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
tail -1 gpurun_out/${R}_plain.log
case $PASS in
launches)
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1 ;;
dram)
    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none -k regex:gemm --csv --log-file gpurun_out/${R}_gemm_dram.csv $CMD \
        > gpurun_out/${R}_ncu_dram.log 2>&1 ;;
full)
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_pair_bf16_kernel \
        -s 40 -c 3 -f -o gpurun_out/${R}_gemm_full $CMD > gpurun_out/${R}_ncu_full.log 2>&1 ;;
esac
echo "ncu $PASS rc=$?"
