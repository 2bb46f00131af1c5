#!/bin/bash
# A/B of two builds of the library on the same box: alternate the driver's bench command between them
ALT=${1:-climate2weather_b200/libc2w_b200_hint.so}
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('base', d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
  C2W_LIB=$ALT python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('alt ', d['ms_per_step'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"
done
