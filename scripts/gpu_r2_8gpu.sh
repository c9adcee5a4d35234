#!/bin/bash
# round-2 8-GPU call: BASELINE configs 3 (v2), 4 (v3, 16 / GPU) and 5 (discrete + causal, 32 / GPU) data-parallel over one box.
# Usage: gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_r2_8gpu.sh
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 "$@" > $O/n8_$name.json 2> $O/n8_$name.err
  echo "== $name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/n8_$name.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("metric", "value", "ms_per_step", "n_gpus")}, d["config"].get("launch"), d["e2e"]["value"], d["clocks"])
except Exception as e:
    print("parse:", e); print(open("gpurun_out/n8_$name.err").read()[-800:])
PY
}
run v2 29531 --steps 8 --warmup 4
run discrete 29532 --config discrete --batch 32 --steps 6 --warmup 3
run v3 29533 --config v3 --batch 16 --steps 3 --warmup 3
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $O/n8_gpus.txt
du -sh gpurun_out
