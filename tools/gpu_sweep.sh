#!/bin/bash
# Quick GPU session: parity tests, K1 timing sweep, chunk-size sweep of the bench step.
TAG=${1:-s01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/test_$TAG.log 2>&1; echo "pytest exit=$?"; tail -3 $OUT/test_$TAG.log
timeout 300 python tools/bringup_conv.py --time > $OUT/bringup_$TAG.log 2>&1; tail -9 $OUT/bringup_$TAG.log
for ch in ${CHUNKS:-8 16 32 64 156}; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --chunk $ch 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('chunk', d['config']['chunk_windows'], 'ms/step', d['ms_per_step'], 'fps', d['value'], 'k1_ms', r['k1_ms_per_step'], 'other_ms', r['other_fwd_kernels_ms_per_step'], 'k1_tf', r['achieved'])"
done | tee $OUT/sweep_$TAG.log
