#!/bin/bash
# L2 traffic of one GEMM shape under the pair kernel and the 4-CTA-cluster (B multicast) kernel.
TAG=$1
M="lts__t_bytes.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max"
for CASE in vis_fc_wgrad vis_qkv_fwd; do
  for Q in 0 1; do
    B200CLIP_GEMM_QUAD=$Q ncu --metrics $M --clock-control none -k regex:gemm -s 2 -c 1 --csv \
        --log-file gpurun_out/${TAG}_${CASE}_quad${Q}.csv python tools/gemm_one.py $CASE > gpurun_out/${TAG}_${CASE}_quad${Q}.log 2>&1
    echo "$CASE quad=$Q rc=$?"
  done
done
