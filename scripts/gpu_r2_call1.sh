#!/bin/bash
# round-2 GPU call 1: tests, per-layer table, ncu --set full of the HBM-bound launches, default bench with the cuDNN arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "exit $rc"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | cut -c1-200
if [ $rc -ne 0 ]; then echo "== pytest (dense PQMF)"; RAVE_PQMF_DENSE=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_dense.log 2>&1; echo "exit $?"; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu_dense.log | cut -c1-200; fi
grep -E "gradient (rel|cos)" gpurun_out/pytest_gpu.log
echo "== layer table"; REPS=20 timeout 300 python scripts/ncu_layers.py 2>&1 | tee gpurun_out/r2_layers_table.txt
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc|wgrad_tc|pqmf' -o gpurun_out/r2_ncu_layers -f python scripts/ncu_layers.py > gpurun_out/ncu_layers.log 2>&1; echo "exit $?"
ncu -i gpurun_out/r2_ncu_layers.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size > gpurun_out/r2_ncu_layers_raw.csv 2>/dev/null
echo "== bench"; timeout 900 python bench.py --cudnn-baseline > gpurun_out/bench_r2_call1.json 2> gpurun_out/bench_r2_call1.err; echo "exit $?"; cut -c1-1500 gpurun_out/bench_r2_call1.json
