#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_engine.py -q -m gpu --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -4 gpurun_out/pytest_all.log | cut -c1-200
echo "== bench N=1"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "exit $?"
tail -1 gpurun_out/bench_bf16.log | cut -c1-330; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
echo "== bench N=1 no wgrad stream"; RAVE_WGRAD_STREAM=0 timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16_nows.log 2> gpurun_out/bench_bf16_nows.err; echo "exit $?"
tail -1 gpurun_out/bench_bf16_nows.log | cut -c1-330
