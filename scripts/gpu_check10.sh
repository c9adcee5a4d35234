#!/bin/bash
mkdir -p gpurun_out
echo "== halo check"; timeout 300 python scripts/check_halo.py > gpurun_out/check_halo.log 2>&1; echo "exit $?"; cat gpurun_out/check_halo.log | cut -c1-220
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_tc.py -q -m gpu --timeout 600 > gpurun_out/pytest_sel.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sel.log; tail -6 gpurun_out/pytest_sel.log | cut -c1-250
echo "== bench N=1 halo off"; RAVE_TC_HALO=0 timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16_nohalo.log 2> gpurun_out/bench_bf16_nohalo.err; echo "exit $?"
tail -1 gpurun_out/bench_bf16_nohalo.log | cut -c1-330; tail -3 gpurun_out/bench_bf16_nohalo.err | cut -c1-300
echo "== bench N=1 halo on"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "exit $?"
tail -1 gpurun_out/bench_bf16.log | cut -c1-330; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
echo "== trace"; timeout 600 python scripts/trace_step.py > gpurun_out/trace_step.txt 2> gpurun_out/trace_step.err; echo "exit $?"; head -30 gpurun_out/trace_step.txt | cut -c1-160; tail -3 gpurun_out/trace_step.err | cut -c1-300
echo "== layers"; timeout 600 python scripts/profile_layers.py > gpurun_out/layers.txt 2> gpurun_out/layers.err; echo "exit $?"; head -12 gpurun_out/layers.txt | cut -c1-200; tail -3 gpurun_out/layers.err | cut -c1-300
echo "== ncu full conv_tc (msd 384->768)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 1 -o gpurun_out/prof_tc3 -f python scripts/prof_conv_tc.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-200
