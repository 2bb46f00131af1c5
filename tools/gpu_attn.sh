#!/bin/bash
for qb in 16 32 64; do
  echo "== C2W_ATTN_QB=$qb"
  C2W_ATTN_QB=$qb timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step', d['ms_per_step'], 'k1_ms', r['k1_ms_per_step'], 'other_ms', r['other_fwd_kernels_ms_per_step'], 'clocks', d['clocks'])"
done
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attention" 2>&1 | tail -2
