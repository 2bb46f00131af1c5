#!/bin/bash
# N-GPU session: config 3 (L=720 time-sharded, strong scaling) and config 4 (16-member ensemble of a year) bench lines
TAG=${1:-r02n}
N=${2:-8}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}_$TAG.log 2>&1; echo "bench n$N exit=$?"; grep '"metric"' $OUT/bench_n${N}_$TAG.log | cut -c1-300; grep -o '"strong_scaling.*' $OUT/bench_n${N}_$TAG.log | cut -c1-700
if [ "$N" = 8 ]; then
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config 4 --steps 3 --warmup 3 > $OUT/bench_config4_n${N}_$TAG.log 2>&1; echo "config4 n$N exit=$?"; grep '"metric"' $OUT/bench_config4_n${N}_$TAG.log | cut -c1-1200; tail -3 $OUT/bench_config4_n${N}_$TAG.log | grep -v metric | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --mode train --steps 5 --warmup 3 > $OUT/bench_train_n${N}_$TAG.log 2>&1; echo "train n$N exit=$?"; grep '"metric"' $OUT/bench_train_n${N}_$TAG.log | cut -c1-300
else
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > $OUT/test_multi_n${N}_$TAG.log 2>&1; echo "multi pytest exit=$?"; grep -E "passed|failed|skipped|^E  " $OUT/test_multi_n${N}_$TAG.log | tail -5
fi
