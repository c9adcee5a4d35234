#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt
echo "== tc fwd tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --timeout 120 -k "tc_vs_oracle" > gpurun_out/pytest_tc_fwd.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc_fwd.log; tail -30 gpurun_out/pytest_tc_fwd.log
echo "== tc wgrad tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -q -m gpu --timeout 120 -k "wgrad" > gpurun_out/pytest_tc_wg.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tc_wg.log; tail -30 gpurun_out/pytest_tc_wg.log
echo "== bench"; timeout 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.log
