#!/bin/bash
# 2-GPU session: multi-GPU parity tests, config 3 (L=720 time-sharded) bench line, DDP training line
TAG=${1:-r02m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_multi.py -q -s > $OUT/test_multi_$TAG.log 2>&1; echo "multi pytest exit=$?"; grep -E "passed|failed|skipped|^E  " $OUT/test_multi_$TAG.log | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n${N}_$TAG.log 2>&1; echo "bench n$N exit=$?"; grep '"metric"' $OUT/bench_n${N}_$TAG.log | cut -c1-2500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --mode train --steps 5 --warmup 3 > $OUT/bench_train_n${N}_$TAG.log 2>&1; echo "train n$N exit=$?"; grep '"metric"' $OUT/bench_train_n${N}_$TAG.log | cut -c1-900; tail -5 $OUT/bench_train_n${N}_$TAG.log | grep -v metric | cut -c1-300
