#!/bin/bash
# A/B of an environment switch on the bench step: bash tools/gpu_ab.sh VAR [v0 v1 ...]
VAR=${1:-C2W_NO_AR}; shift
VALS=${@:-0 1}
OUT=gpurun_out; mkdir -p $OUT
for v in $VALS; do
  env $VAR=$v timeout 300 python bench.py --steps 16 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$VAR=$v', 'ms/step', d['ms_per_step'], 'fps', d['value'], 'k1_ms', r['k1_ms_per_step'], 'other_ms', r['other_fwd_kernels_ms_per_step'], 'k1_tf', r['achieved'], d['clocks']['sm_mhz'])"
done 2>&1 | tee $OUT/ab_$VAR.log
