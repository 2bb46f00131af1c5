#!/bin/bash
TAG=${1:-ex}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $OUT/launches_exact_$TAG.csv python bench.py --exact-grad --profile --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_launches_exact_$TAG.log 2>&1; echo "ncu exit=$?"
python tools/ncu_launches.py $OUT/launches_exact_$TAG.csv > $OUT/launch_shares_exact_$TAG.txt 2>&1; head -40 $OUT/launch_shares_exact_$TAG.txt
