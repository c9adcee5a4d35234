#!/bin/bash
# last call of the round: fused feature tap + time stack, forward and backward (v3 MRD) -- GPU suite in one process, v3 quick bench
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 200 python -m pytest tests -x -q -m gpu > $O/c32_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c32_tests.log | cut -c1-400
timeout 100 python bench.py --config v3 --batch 16 --steps 8 --warmup 3 --quick > $O/c32_v3.json 2> $O/c32_v3.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c32_v3.json").read().strip().splitlines()[-1])
    print("v3", round(d["ms_per_step"], 3))
except Exception as e:
    print("v3 parse:", e); print(open("gpurun_out/c32_v3.err").read()[-500:])
PY
