#!/bin/bash
# Round-2 GPU session (one B200): parity tests (incl. full-length + bit-exact compose), bench line, ncu launch list with
# DRAM bytes, ncu --set full of the residual+LayerNorm K1 variants.  Usage: gpurun -- bash tools/gpu_r02.sh <tag> [quick]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.csv
timeout 1500 python -m pytest tests -m gpu -q -s -x --deselect tests/test_gpu_library_bar.py > $OUT/test_$TAG.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" $OUT/test_$TAG.log | tail -3
grep -E "^\[|x after|first guided|final \(|worst frame" $OUT/test_$TAG.log | head -60
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_$TAG.log | cut -c1-1500
C2W_PDL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_nopdl_$TAG.log 2>&1; echo "bench nopdl exit=$?"; tail -1 $OUT/bench_nopdl_$TAG.log | cut -c1-400
[ "$2" = quick ] && exit 0
timeout 600 python bench.py --steps 6 --warmup 3 --exact-grad --no-cpu > $OUT/bench_exact_$TAG.log 2>&1; echo "bench exact exit=$?"; tail -1 $OUT/bench_exact_$TAG.log | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
    --log-file $OUT/launches_$TAG.csv python bench.py --profile --steps 1 --warmup 1 > $OUT/ncu_launches_$TAG.log 2>&1; echo "ncu launches exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tcgen05 -c 8 -f -o $OUT/prof_conv_$TAG \
    python tools/bringup_conv.py --ncu-variants > $OUT/ncu_conv_$TAG.log 2>&1; echo "ncu conv exit=$?"
