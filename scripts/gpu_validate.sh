#!/bin/bash
# One GPU-box validation pass: all GPU tests, smoke(), the bench line.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_validate.sh'
mkdir -p gpurun_out
echo "== tests (all gpu)"; timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log | cut -c1-200
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/smoke.log | cut -c1-200
echo "== bench N=1"; timeout 1200 python bench.py --steps 8 --warmup 4 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "exit $?"
tail -1 gpurun_out/bench_default.log | cut -c1-3000; tail -3 gpurun_out/bench_default.err | cut -c1-300
