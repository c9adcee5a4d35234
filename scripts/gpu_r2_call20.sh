#!/bin/bash
# streaming re-check, channel-last Descript discriminator tests, v3 bench + trace
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_descript.py tests/test_gpu_parity.py -m gpu -q -k "descript or streaming or time_stack or v3 or activation or am_tanh" > gpurun_out/c20_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/c20_tests.log
tail -5 gpurun_out/c20_tests.log
timeout 400 python bench.py --config v3 --batch 16 --steps 5 --warmup 3 --quick > gpurun_out/c20_bench_v3.json 2> gpurun_out/c20_bench_v3.err
tail -c 600 gpurun_out/c20_bench_v3.json
timeout 300 python scripts/trace_step_config.py v3 16 > gpurun_out/c20_trace_v3.txt 2>&1
head -40 gpurun_out/c20_trace_v3.txt
