#!/bin/bash
# ncu --set full (source-level) capture of one long-sequence attention forward launch (ViT-B/16, 1024 images, S = 197).
TAG=$1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fwd_long -c 1 -f \
    -o gpurun_out/${TAG}_attn_fwd_long python tools/profile_infer.py 3 > gpurun_out/${TAG}_ncu_long.log 2>&1
echo "ncu long rc=$?"
