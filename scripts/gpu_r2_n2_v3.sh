#!/bin/bash
# BASELINE config 4 (v3, 16 / GPU) data-parallel over 2 GPUs of one box, current tree.  Usage: gpurun --gpus 2 -- bash scripts/gpu_r2_n2_v3.sh
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config v3 --batch 16 --steps 8 --warmup 3 --quick > gpurun_out/n2_v3.json 2> gpurun_out/n2_v3.err
echo "exit $?"; tail -c 700 gpurun_out/n2_v3.json; tail -3 gpurun_out/n2_v3.err
