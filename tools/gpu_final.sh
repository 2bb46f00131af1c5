OUT=gpurun_out; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 16 --warmup 3 > $OUT/bench_r01g.log 2>&1; echo "bench exit=$?"; tail -1 $OUT/bench_r01g.log | cut -c1-300
timeout 200 python bench.py --steps 4 --warmup 3 --exact-grad --no-cpu > $OUT/bench_exact_r01g.log 2>&1; tail -1 $OUT/bench_exact_r01g.log | cut -c1-200
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_r01g.csv python bench.py --profile --steps 1 --warmup 1 > $OUT/ncu_launches_r01g.log 2>&1; echo "ncu exit=$?"
