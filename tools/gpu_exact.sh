#!/bin/bash
# exact_grad iteration: parity + full-length tests, the exact-grad bench line
TAG=${1:-ex}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_length.py -m gpu -q -s > $OUT/test_exact_$TAG.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" $OUT/test_exact_$TAG.log | tail -3
grep -E "^FAILED|^ERROR|exact" $OUT/test_exact_$TAG.log | head -20
timeout 600 python bench.py --exact-grad --steps 10 --warmup 3 --no-cpu > $OUT/bench_exact_grad_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_exact_grad_$TAG.log | cut -c1-1200
