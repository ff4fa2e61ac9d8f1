#!/bin/bash
# Same-box A/B of the 1-GPU headline step under a few switches.  usage: tools/ab_step.sh <tag>
TAG=$1
run() {
    name=$1; shift
    env "$@" B200CLIP_BENCH_VARIANTS=0 BENCH_SKIP_CPU=1 timeout 300 python bench.py --steps 10 --warmup 3 \
        > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${name}.json"))
    print("${name}", round(d["value"]), round(d["ms_per_step"], 3), "gemm", round(d["roofline"]["achieved"]), d["clocks"]["sm_mhz"])
except Exception as e:
    print("${name} failed", e)
PY
}
run default X=1
run d8off B200CLIP_QGELU_D8=0
run d8off_tmaoff B200CLIP_QGELU_D8=0 B200CLIP_EPI_TMA=0
run default2 X=1
