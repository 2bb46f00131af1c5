#!/bin/bash
# ncu session: launch lists (time + DRAM bytes) of a sampling step and a training step, full capture of the K1 variants
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --profile --steps 1 --warmup 1 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --mode train --profile --steps 1 --warmup 1 --batch 128 > $OUT/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -c 8 -f -o $OUT/prof_conv_$TAG \
    python tools/bringup_conv.py --ncu-variants > $OUT/ncu_conv_$TAG.log 2>&1; echo "ncu conv exit=$?"
ls -la $OUT | tail -8
