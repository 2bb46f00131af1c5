#!/bin/bash
# A/B of an environment switch on the K1 sweep and the bench step: bash tools/gpu_ab.sh VAR
VAR=${1:-C2W_L2_PREFETCH}
OUT=gpurun_out; mkdir -p $OUT
for v in 0 1; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 300 python tools/bringup_conv.py --time 2>&1 | grep shape | cut -c1-120 | head -6
  env $VAR=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step', d['ms_per_step'], 'fps', d['value'], 'k1_ms', r['k1_ms_per_step'], 'other_ms', r['other_fwd_kernels_ms_per_step'], 'k1_tf', r['achieved'])"
done 2>&1 | tee $OUT/ab_$VAR.log
