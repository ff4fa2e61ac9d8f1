#!/bin/bash
# Validation ladder for the packed text tower before it may become the default (run under gpurun).
#   tools/validate_pack.sh 1      one GPU : whole GPU suite + bench with B200CLIP_PACK_TEXT=1, then =2 (graphs)
#   tools/validate_pack.sh 2      two GPUs: tools/dist_check.py + bench --gpus 2 under both settings
# Every stage runs under its own timeout; logs in gpurun_out/pack_*.log.
set -u
N=${1:-1}
for MODE in 1 2; do
    export B200CLIP_PACK_TEXT=$MODE
    if [ "$N" = "1" ]; then
        timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pack_${MODE}_tests.log 2>&1
        echo "PACK_TEXT=$MODE pytest rc=$? : $(tail -1 gpurun_out/pack_${MODE}_tests.log)"
        timeout 200 python bench.py > gpurun_out/pack_${MODE}_bench.json 2> gpurun_out/pack_${MODE}_bench.err
        echo "PACK_TEXT=$MODE bench rc=$? : $(cut -c1-200 gpurun_out/pack_${MODE}_bench.json)"
    else
        timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
            --master-port 2951$MODE tools/dist_check.py > gpurun_out/pack_${MODE}_dist.log 2>&1
        echo "PACK_TEXT=$MODE dist_check rc=$? : $(grep -a DIST_CHECK gpurun_out/pack_${MODE}_dist.log | cut -c1-220)"
        timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
            --master-port 2952$MODE bench.py --gpus $N --steps 10 --warmup 3 \
            > gpurun_out/pack_${MODE}_bench_n$N.json 2> gpurun_out/pack_${MODE}_bench_n$N.err
        echo "PACK_TEXT=$MODE bench --gpus $N rc=$? : $(cut -c1-200 gpurun_out/pack_${MODE}_bench_n$N.json)"
    fi
done
