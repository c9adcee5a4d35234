#!/bin/bash
mkdir -p gpurun_out
echo "== 2cta check"; RAVE_TC_2CTA=1 timeout 300 python scripts/check_2cta.py > gpurun_out/check_2cta.log 2>&1; echo "exit $?" >> gpurun_out/check_2cta.log; tail -12 gpurun_out/check_2cta.log | cut -c1-300
echo "== gpu tests (tc + engine)"; timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py -q -m gpu --timeout 600 > gpurun_out/pytest_tc_engine.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc_engine.log; tail -12 gpurun_out/pytest_tc_engine.log | cut -c1-250
echo "== bench bf16 graphs"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "bench exit $?" >> gpurun_out/bench_bf16.err
tail -1 gpurun_out/bench_bf16.log | cut -c1-700; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
echo "== bench bf16 graphs + 2cta"; RAVE_TC_2CTA=1 timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16_2cta.log 2> gpurun_out/bench_bf16_2cta.err; echo "bench exit $?" >> gpurun_out/bench_bf16_2cta.err
tail -1 gpurun_out/bench_bf16_2cta.log | cut -c1-700; tail -3 gpurun_out/bench_bf16_2cta.err | cut -c1-300
