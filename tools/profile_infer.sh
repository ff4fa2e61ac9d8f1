#!/bin/bash
# Launch list (gpu__time_duration.sum per launch) of one no-grad forward of BASELINE config 3 and 4.
# usage (under gpurun): tools/profile_infer.sh <tag>
TAG=$1
for C in 3 4; do
    python tools/profile_infer.py $C > gpurun_out/${TAG}_cfg${C}_plain.log 2>&1 || { echo "plain cfg$C failed"; tail -3 gpurun_out/${TAG}_cfg${C}_plain.log; continue; }
    tail -1 gpurun_out/${TAG}_cfg${C}_plain.log
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/${TAG}_cfg${C}_launches.csv python tools/profile_infer.py $C > gpurun_out/${TAG}_cfg${C}_ncu.log 2>&1
    echo "ncu cfg$C rc=$?"
done
