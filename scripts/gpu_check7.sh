#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log | cut -c1-250
echo "== bench bf16 graphs"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "bench exit $?" >> gpurun_out/bench_bf16.err
tail -1 gpurun_out/bench_bf16.log | cut -c1-1800; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
echo "== ncu launch list (bf16 eager)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 3000 --csv --log-file gpurun_out/launches_bf16_v4.csv python bench.py --steps 4 --warmup 4 --precision bf16 --no-cpu-baseline --no-graphs > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
echo "== kernel shapes"; timeout 300 python scripts/prof_conv_tc.py 2>&1 | tee gpurun_out/shapes.log
