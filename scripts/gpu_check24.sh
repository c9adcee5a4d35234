#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "== tests (all gpu)"; timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --durations=5 > gpurun_out/pytest_all.log 2>&1; echo "exit $?" >> gpurun_out/pytest_all.log; tail -12 gpurun_out/pytest_all.log | cut -c1-200
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?"; tail -3 gpurun_out/smoke.log | cut -c1-200
echo "== bench N=1 (full: cpu baseline + cudnn baseline)"; timeout 1200 python bench.py --steps 8 --warmup 4 --precision bf16 --cudnn-baseline > gpurun_out/bench_bf16_full.log 2> gpurun_out/bench_bf16_full.err; echo "exit $?"
tail -1 gpurun_out/bench_bf16_full.log | cut -c1-3000; tail -3 gpurun_out/bench_bf16_full.err | cut -c1-300
echo "== ncu full conv_tc2<256,64>"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -o gpurun_out/prof_tc2_r1 -f python scripts/prof_conv_tc.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_full.log | cut -c1-200
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 1 --precision bf16 --no-graphs --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
