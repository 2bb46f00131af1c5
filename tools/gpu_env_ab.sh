#!/bin/bash
# A/B of an environment switch on the same box: alternate the driver's bench command with and without it
VAR=${1:-C2W_A_PREFETCH=1}
for i in 1 2 3; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base', d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
  env $VAR python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$VAR', d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done
