#!/bin/bash
# round-2 GPU call 1: tests, per-layer table, ncu --set full of the HBM-bound launches, default bench with the cuDNN arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== pytest (one process per group)"; bash scripts/gpu_tests_isolated.sh
echo "== layer table"; REPS=20 timeout 300 python scripts/ncu_layers.py 2>&1 | tee gpurun_out/r2_layers_table.txt
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc|pqmf|dilated_unit' -o gpurun_out/r2_ncu_layers -f python scripts/ncu_layers.py > gpurun_out/ncu_layers.log 2>&1; echo "exit $?"
ncu -i gpurun_out/r2_ncu_layers.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size > gpurun_out/r2_ncu_layers_raw.csv 2>/dev/null
echo "== bench"; timeout 900 python bench.py --cudnn-baseline > gpurun_out/bench_r2_call1.json 2> gpurun_out/bench_r2_call1.err; echo "exit $?"; cut -c1-1500 gpurun_out/bench_r2_call1.json
