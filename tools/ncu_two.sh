#!/bin/bash
# ncu --set full (source-level) captures of one attention-backward launch and one fat-epilogue GEMM launch.
# usage (under gpurun): tools/ncu_two.sh <tag>
TAG=$1
python tools/ncu_micro.py --only="vis attn" > gpurun_out/${TAG}_plain_attn.log 2>&1 || { echo plain attn failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -s 1 -c 1 -f -o gpurun_out/${TAG}_attn_bwd \
    python tools/ncu_micro.py --only="vis attn" > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
python tools/gemm_one.py vis_fc_fwd 6400 > gpurun_out/${TAG}_plain_gemm.log 2>&1 || { echo plain gemm failed; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:gemm -s 2 -c 1 -f -o gpurun_out/${TAG}_fc_fwd_6400 \
    python tools/gemm_one.py vis_fc_fwd 6400 > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
