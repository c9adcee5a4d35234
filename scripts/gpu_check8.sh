#!/bin/bash
mkdir -p gpurun_out
echo "== tc + engine + parity spectral tests"; timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_engine.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "tc or engine or spectral or bf16 or fused" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?" >> gpurun_out/pytest_sel.log; tail -6 gpurun_out/pytest_sel.log | cut -c1-250
echo "== bench bf16 graphs + cudnn baseline"; timeout 1200 python bench.py --steps 8 --warmup 4 --precision bf16 --cudnn-baseline > gpurun_out/bench_bf16.log 2> gpurun_out/bench_bf16.err; echo "bench exit $?" >> gpurun_out/bench_bf16.err
tail -1 gpurun_out/bench_bf16.log | cut -c1-2600; tail -3 gpurun_out/bench_bf16.err | cut -c1-300
echo "== kernel shapes"; timeout 300 python scripts/prof_conv_tc.py 2>&1 | tee gpurun_out/shapes.log
