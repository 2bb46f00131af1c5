#!/bin/bash
# quick regression session: GPU tests (without the full-length file) + bench with and without PDL
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -s --deselect tests/test_gpu_library_bar.py --ignore tests/test_gpu_full_length.py > $OUT/test_$TAG.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" $OUT/test_$TAG.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/test_$TAG.log | head -20
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_$TAG.log | cut -c1-1800
C2W_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_nopdl_$TAG.log 2>&1; echo "bench nopdl exit=$?"; tail -1 $OUT/bench_nopdl_$TAG.log | cut -c1-300
