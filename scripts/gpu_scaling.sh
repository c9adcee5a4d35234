#!/bin/bash
# N=1 vs N=2 (same box) bench lines + the reference arm.  Usage: gpurun --gpus 2 --timeout 1500 -- bash scripts/gpu_scaling.sh
mkdir -p gpurun_out
echo "== bench N=1"; timeout 600 python bench.py --steps 8 --warmup 4 --precision bf16 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err; echo "exit $?"
tail -1 gpurun_out/bench_n1.log | cut -c1-400; tail -2 gpurun_out/bench_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.log').read().strip().splitlines()[-1]); print("clocks", d["clocks"], "e2e", d["e2e"]["value"])
PY
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 4 --precision bf16 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "exit $?"
tail -1 gpurun_out/bench_n2.log | cut -c1-400; tail -3 gpurun_out/bench_n2.err | cut -c1-300
echo "== reference arm (N=1)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "exit $?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
