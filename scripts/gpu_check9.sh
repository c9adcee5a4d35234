#!/bin/bash
mkdir -p gpurun_out
echo "== engine tests"; timeout 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_tc.py -q -m gpu --timeout 600 > gpurun_out/pytest_sel.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sel.log; tail -4 gpurun_out/pytest_sel.log | cut -c1-250
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 4 --precision bf16 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.log | cut -c1-900; tail -5 gpurun_out/bench_n2.err | cut -c1-400
echo "== bench N=1 + cudnn baseline"; timeout 900 python bench.py --steps 8 --warmup 4 --precision bf16 --cudnn-baseline --no-cpu-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "exit $?" >> gpurun_out/bench_bf16.err
tail -1 gpurun_out/bench_bf16.log | cut -c1-3000; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
