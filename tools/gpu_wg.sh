#!/bin/bash
# K10 iteration: wgrad + training tests, the K10 sweep with and without the bias gradient, the training line
TAG=${1:-wg}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_wgrad.py tests/test_gpu_train.py -m gpu -q -x > $OUT/test_wg_$TAG.log 2>&1; echo "pytest exit=$?"
tail -5 $OUT/test_wg_$TAG.log
timeout 300 python tools/bringup_wgrad.py > $OUT/wgrad_sweep_$TAG.log 2>&1; echo "sweep exit=$?"; cat $OUT/wgrad_sweep_$TAG.log
timeout 300 python tools/bringup_wgrad.py --no-bias > $OUT/wgrad_sweep_nobias_$TAG.log 2>&1; echo "sweep exit=$?"; cat $OUT/wgrad_sweep_nobias_$TAG.log
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > $OUT/bench_train_n1_$TAG.log 2>&1; echo "train exit=$?"; tail -1 $OUT/bench_train_n1_$TAG.log | cut -c1-300
