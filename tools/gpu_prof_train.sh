#!/bin/bash
TAG=${1:-tr}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --mode train --profile --steps 1 --warmup 1 --batch 128 > $OUT/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches exit=$?"
python tools/ncu_launches.py $OUT/launches_train_$TAG.csv > $OUT/launch_shares_train_$TAG.txt 2>&1; head -45 $OUT/launch_shares_train_$TAG.txt
