#!/bin/bash
# Descript feature taps + fake-rows-only backward: tests, v3 bench + trace
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_descript.py tests/test_gpu_parity.py -m gpu -q -k "descript or streaming or time_stack or v3 or am_tanh or leaky or training_step" > gpurun_out/c21_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c21_tests.log
tail -5 gpurun_out/c21_tests.log
timeout 400 python bench.py --config v3 --batch 16 --steps 6 --warmup 3 --quick > gpurun_out/c21_bench_v3.json 2> gpurun_out/c21_bench_v3.err
tail -c 400 gpurun_out/c21_bench_v3.json
timeout 300 python scripts/trace_step_config.py v3 16 > gpurun_out/c21_trace_v3.txt 2>&1
head -30 gpurun_out/c21_trace_v3.txt
