#!/bin/bash
# Multi-GPU session (gpurun --gpus N): sharded-sampling parity + weak-scaling bench lines.
N=${1:-2}; TAG=${2:-m01}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/test_multi_$TAG.log 2>&1; echo "pytest exit=$?"; tail -5 $OUT/test_multi_$TAG.log
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' > $OUT/bench_${TAG}_n1.log
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 8 --warmup 3 --no-cpu 2>&1 | grep '^{' > $OUT/bench_${TAG}_n$n.log; echo "bench n=$n exit=$?"
  fi
done
if [ $N -ge 8 ]; then  # BASELINE.json config 3: one month (720 frames) time-sharded over 8 GPUs
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu --frames 720 2>&1 | grep '^{' > $OUT/bench_${TAG}_L720_n8.log; echo "bench L=720 n=8 exit=$?"
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*_n[0-9].log')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'gpus',d['n_gpus'],'frames',d['config']['frames'],'fps',d['value'],'ms/step',d['ms_per_step'],'e2e',d.get('e2e',{}).get('value'))
    except Exception as e: print(f,'ERR',e)
PY
