#!/bin/bash
# end-of-round check on one GPU: whole GPU suite, smoke, the driver's bench line, the training line, a training launch list
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/test_gpu_all_$TAG.log 2>&1; echo "pytest exit=$?"
grep -E "passed|failed|error" $OUT/test_gpu_all_$TAG.log | tail -3
grep -E "^FAILED|^ERROR" $OUT/test_gpu_all_$TAG.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit=$?"; tail -2 $OUT/smoke_$TAG.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_$TAG.log | cut -c1-1500
timeout 600 python bench.py --mode train --steps 5 --warmup 3 > $OUT/bench_train_n1_$TAG.log 2>&1; echo "train exit=$?"; tail -1 $OUT/bench_train_n1_$TAG.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv \
    --log-file $OUT/launches_train_$TAG.csv python bench.py --mode train --profile --steps 1 --warmup 1 --batch 128 > $OUT/ncu_launches_train_$TAG.log 2>&1; echo "ncu train launches exit=$?"
python tools/ncu_launches.py $OUT/launches_train_$TAG.csv > $OUT/launch_shares_train_$TAG.txt 2>&1; head -30 $OUT/launch_shares_train_$TAG.txt
