#!/bin/bash
# One GPU session: parity tests, K1 timing sweep, bench line (+ exact-grad line), ncu launch list + full capture of K1.
# Usage (via gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/test_$TAG.log 2>&1; echo "pytest exit=$?"; grep -E "rel-err|passed|failed" $OUT/test_$TAG.log | tail -12
timeout 300 python tools/bringup_conv.py --time > $OUT/bringup_$TAG.log 2>&1; grep shape $OUT/bringup_$TAG.log
timeout 600 python bench.py --steps 16 --warmup 3 > $OUT/bench_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_$TAG.log
timeout 600 python bench.py --steps 4 --warmup 3 --exact-grad --no-cpu > $OUT/bench_exact_$TAG.log 2>&1; echo "bench exact exit=$?"; tail -1 $OUT/bench_exact_$TAG.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --profile --steps 1 --warmup 1 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit=$?"
[ -n "$SKIP_NCU_FULL" ] || { timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -s 4 -c 2 -f -o $OUT/prof_conv_$TAG \
    python tools/bringup_conv.py --only-g2 > $OUT/ncu_conv_$TAG.log 2>&1; echo "ncu conv exit=$?"; }
