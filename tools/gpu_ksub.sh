#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/bringup_conv.py --time 2>&1 | grep shape | head -8 > $OUT/ksub2.log; cat $OUT/ksub2.log
# rebuild with one K sub-block per stage for CTA pairs
sed -i 's/#define C2W_KSUB_CG2 2/#define C2W_KSUB_CG2 1/' climate2weather_b200/csrc/conv_tcgen05.cuh
python -m climate2weather_b200.build > /dev/null 2>&1; echo "build exit=$?"
timeout 300 python tools/bringup_conv.py --time 2>&1 | grep shape | head -8 > $OUT/ksub1.log; cat $OUT/ksub1.log
