#!/bin/bash
# One gpurun call = a list of stages, each under its own timeout, each logging to gpurun_out/<tag>_<stage>.log.
#   tools/gpu_call.sh <tag> stage1 stage2 ...      stages: tests | bench | bench_nopack | gemm | gemm128 | micro | infer
set -u
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
for STAGE in "$@"; do
    T0=$(date +%s)
    case $STAGE in
        tests)        timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1 ;;
        attn)         timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "attention or attn" > gpurun_out/${TAG}_attn.log 2>&1 ; tail -5 gpurun_out/${TAG}_attn.log ;;
        graph128_nomin) B200CLIP_MIN_GRID=0 timeout 400 python tools/graph_bench.py 128 > gpurun_out/${TAG}_graph128_nomin.txt 2>&1 ;;
        gemmtests)    timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "gemm" > gpurun_out/${TAG}_gemmtests.log 2>&1 ; tail -5 gpurun_out/${TAG}_gemmtests.log ;;
        micro_notma)  B200CLIP_EPI_TMA=0 timeout 300 python tools/ncu_micro.py --time --only=fc_fwd > gpurun_out/${TAG}_micro_notma.txt 2>&1 ;;
        gemm_noquad)  B200CLIP_GEMM_QUAD=0 timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/${TAG}_gemm1024_noquad.txt 2>&1 ;;
        gemmquick)    timeout -s KILL 150 python -m pytest tests -m gpu -q -x --timeout 60 -k "gemm" > gpurun_out/${TAG}_gemmtests.log 2>&1 ; tail -5 gpurun_out/${TAG}_gemmtests.log ;;
        tests_all)    timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_tests.log 2>&1 ;;
        bench)        timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
        bench_nv)     B200CLIP_BENCH_VARIANTS=0 timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ;;
        bench_nopack) B200CLIP_PACK_TEXT=0 B200CLIP_BENCH_VARIANTS=0 timeout 400 python bench.py --steps 5 > gpurun_out/${TAG}_bench_nopack.json 2> gpurun_out/${TAG}_bench_nopack.err ;;
        bench128)     B200CLIP_BENCH_VARIANTS=0 BENCH_GLOBAL_BATCH=128 timeout 400 python bench.py --steps 20 > gpurun_out/${TAG}_bench128.json 2> gpurun_out/${TAG}_bench128.err ;;
        gemm)         timeout 300 python tools/gemm_bench.py 1024 > gpurun_out/${TAG}_gemm1024.txt 2>&1 ;;
        gemm128)      timeout 300 python tools/gemm_bench.py 128 > gpurun_out/${TAG}_gemm128.txt 2>&1 ;;
        micro)        timeout 300 python tools/ncu_micro.py --time > gpurun_out/${TAG}_micro.txt 2>&1 ;;
        infer)        timeout 400 python tools/infer_bench.py > gpurun_out/${TAG}_infer.txt 2>&1 ;;
        launches128)  PAIRS=128 bash tools/collect_profiles.sh launches ${TAG}_b128 > gpurun_out/${TAG}_launches128.log 2>&1 ;;
        launches)     PAIRS=1024 bash tools/collect_profiles.sh launches ${TAG}_b1024 > gpurun_out/${TAG}_launches1024.log 2>&1 ;;
        dist2)        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py > gpurun_out/${TAG}_dist2.log 2>&1 ;;
        bench2)       B200CLIP_BENCH_VARIANTS=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err ;;
        graph128)     timeout 400 python tools/graph_bench.py 128 > gpurun_out/${TAG}_graph128.txt 2>&1 ;;
        graph1024)    timeout 400 python tools/graph_bench.py 1024 > gpurun_out/${TAG}_graph1024.txt 2>&1 ;;
        cfg1)         timeout 300 python bench.py --config 1 --steps 50 > gpurun_out/${TAG}_cfg1.json 2> gpurun_out/${TAG}_cfg1.err ;;
        cfg3)         timeout 300 python bench.py --config 3 --steps 3 > gpurun_out/${TAG}_cfg3.json 2> gpurun_out/${TAG}_cfg3.err ;;
        cfg4)         timeout 300 python bench.py --config 4 --steps 5 > gpurun_out/${TAG}_cfg4.json 2> gpurun_out/${TAG}_cfg4.err ;;
        cfg5small)    BENCH_PAIRS_PER_GPU=64 timeout 400 python bench.py --config 5 --steps 2 > gpurun_out/${TAG}_cfg5small.json 2> gpurun_out/${TAG}_cfg5small.err ;;
        cfg5)         timeout 600 python bench.py --config 5 --steps 3 > gpurun_out/${TAG}_cfg5.json 2> gpurun_out/${TAG}_cfg5.err ;;
        dist8)        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/dist_check.py > gpurun_out/${TAG}_dist8.log 2>&1 ;;
        bench8)       B200CLIP_BENCH_VARIANTS=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench8.json 2> gpurun_out/${TAG}_bench8.err ;;
        bench4)       B200CLIP_BENCH_VARIANTS=0 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench4.json 2> gpurun_out/${TAG}_bench4.err ;;
        cfg5x8)       timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29564 bench.py --gpus 8 --config 5 --steps 3 --warmup 3 > gpurun_out/${TAG}_cfg5x8.json 2> gpurun_out/${TAG}_cfg5x8.err ;;
        smoke)        timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1 ;;
        *)            echo "unknown stage $STAGE" ;;
    esac
    echo "stage $STAGE rc=$? $(( $(date +%s) - T0 ))s"
done
tail -3 gpurun_out/${TAG}_tests.log 2>/dev/null
cut -c1-400 gpurun_out/${TAG}_bench.json 2>/dev/null
