#!/bin/bash
# ncu --set full (source-level) capture of one attention-backward launch (vision shape, 1024 pairs).
TAG=$1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd -c 1 -f -o gpurun_out/${TAG}_attn_bwd \
    python tools/ncu_micro.py --only="vis attn" > gpurun_out/${TAG}_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
