#!/bin/bash
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_tc.py -q -m gpu --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -4 gpurun_out/pytest_all.log | cut -c1-200
for ns in 1 3 8; do
echo "== bench N=1 streams=$ns"; RAVE_DISC_STREAMS=$ns timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_s$ns.log 2> gpurun_out/bench_s$ns.err; echo "exit $?"
tail -1 gpurun_out/bench_s$ns.log | cut -c1-330; tail -2 gpurun_out/bench_s$ns.err | cut -c1-300
done
