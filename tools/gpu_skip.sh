#!/bin/bash
# in-step upper bounds (diagnostics build; outputs are garbage): what would the step cost without weight loads (3),
# activation loads (4), any operand loads (1)?
mkdir -p gpurun_out
for s in ${SKIPS:-0 3 4 1 0}; do
  C2W_LIB=climate2weather_b200/libc2w_b200_diag.so C2W_SKIP_LOADS=$s python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>gpurun_out/skip_err_$s.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('skip=$s', d['ms_per_step'], d['roofline']['k1_ms_per_step'], d['clocks']['sm_mhz'])" || echo "skip=$s failed"
done
