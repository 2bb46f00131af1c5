#!/bin/bash
TAG=${1:-at}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -q -x -k "attention or vjp or exact or train or grad" > $OUT/test_attn_$TAG.log 2>&1; echo "pytest exit=$?"
tail -8 $OUT/test_attn_$TAG.log
timeout 600 python bench.py --exact-grad --steps 10 --warmup 3 --no-cpu --no-e2e > $OUT/bench_exact_grad_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_exact_grad_$TAG.log | cut -c1-250
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > $OUT/bench_train_n1_$TAG.log 2>&1; echo "train exit=$?"; tail -1 $OUT/bench_train_n1_$TAG.log | cut -c1-250
