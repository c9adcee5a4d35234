#!/bin/bash
mkdir -p gpurun_out
echo "== engine tests"; timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu --timeout 300 -x > gpurun_out/pytest_engine.log 2>&1; echo "exit $?" >> gpurun_out/pytest_engine.log; tail -40 gpurun_out/pytest_engine.log | cut -c1-300
echo "== bench bf16"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "bench exit $?" >> gpurun_out/bench_bf16.err
tail -1 gpurun_out/bench_bf16.log | cut -c1-1500; tail -5 gpurun_out/bench_bf16.err | cut -c1-400
echo "== ncu launch list (bf16 bench, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bf16.csv python bench.py --steps 4 --warmup 1 --precision bf16 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
wc -l gpurun_out/launches_bf16.csv
