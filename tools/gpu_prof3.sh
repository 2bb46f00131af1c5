#!/bin/bash
# ncu session at the end of round 2: K10 full capture, sampling launch list (time + DRAM bytes), K1 variants full capture
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -c 4 -f -o $OUT/prof_wgrad_$TAG \
    python tools/bringup_wgrad.py --ncu > $OUT/ncu_wgrad_$TAG.log 2>&1; echo "ncu wgrad exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --profile --steps 1 --warmup 1 --no-cpu --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -c 8 -f -o $OUT/prof_conv_$TAG \
    python tools/bringup_conv.py --ncu-variants > $OUT/ncu_conv_$TAG.log 2>&1; echo "ncu conv exit=$?"
ls -la $OUT | tail -8
