#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 2 (small shapes; K1 at full size is too slow under it)
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
  -k "other_variable or attention_backward or exact_grad_guided or compose_by_window_lists and not 168" > $OUT/sanitizer_new_$TAG.log 2>&1; echo "sanitizer 1 exit=$?"
tail -4 $OUT/sanitizer_new_$TAG.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_wgrad.py tests/test_optim.py tests/test_gpu_train.py -q -m gpu -x \
  -k "not full_arch and not golden and not ema_rate" > $OUT/sanitizer_train_$TAG.log 2>&1; echo "sanitizer 2 exit=$?"
tail -4 $OUT/sanitizer_train_$TAG.log
