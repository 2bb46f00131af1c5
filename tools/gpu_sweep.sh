#!/bin/bash
# Chunk-size sweep of the bench step (windows per UNet launch).
TAG=${1:-s01}
OUT=gpurun_out; mkdir -p $OUT
for ch in ${CHUNKS:-16 26 32 39 52 78 156}; do
  timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --chunk $ch 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('chunk', d['config']['chunk_windows'], 'ms/step', d['ms_per_step'], 'fps', d['value'], 'k1_ms', r['k1_ms_per_step'], 'other_ms', r['other_fwd_kernels_ms_per_step'], 'k1_tf', r['achieved'])"
done | tee $OUT/sweep_$TAG.log
