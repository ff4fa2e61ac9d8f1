#!/bin/bash
# bench.py at N GPUs: peer-memory exchange (default) vs the NCCL all-reduce form.  usage: tools/n8_peer.sh <ngpus> <tag>
N=$1; TAG=$2
run() {
    name=$1; shift
    env "$@" B200CLIP_BENCH_VARIANTS=0 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
        --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 20 --warmup 3 \
        > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}", round(d["value"]), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "loss", d["config"].get("final_loss"), d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name} failed", e)
PY
}
run peer X=1
run allreduce B200CLIP_FEATURE_GATHER=allreduce
