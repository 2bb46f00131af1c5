#!/bin/bash
# Quick GPU session: parity tests, K1 variant timing sweep, one bench line.
TAG=${1:-q01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test_$TAG.log 2>&1; echo "pytest exit=$?"; tail -15 $OUT/test_$TAG.log
timeout 300 python tools/bringup_conv.py --time > $OUT/bringup_$TAG.log 2>&1; tail -16 $OUT/bringup_$TAG.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu > $OUT/bench_$TAG.log 2>&1; echo "bench exit=$?"; tail -2 $OUT/bench_$TAG.log
