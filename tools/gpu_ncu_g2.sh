#!/bin/bash
# ncu --set full capture of K1 on the dominant shape (G2: 128->128 @128^2, 32 windows), default variant.
TAG=${1:-n01}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 4 -c 2 -f -o $OUT/prof_conv_$TAG \
    python tools/bringup_conv.py --only-g2 > $OUT/ncu_conv_$TAG.log 2>&1; echo "ncu conv exit=$?"
tail -3 $OUT/ncu_conv_$TAG.log
