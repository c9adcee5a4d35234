#!/bin/bash
# last call of the round: fused feature tap + time stack (v3 MRD) -- GPU suite in one process, v3 quick bench fused / unfused
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 300 python -m pytest tests -x -q -m gpu > $O/c31_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c31_tests.log | cut -c1-400
b() { name=$1; shift; env "$@" timeout 120 python bench.py --config v3 --batch 16 --steps 8 --warmup 3 --quick > $O/c31_$name.json 2> $O/c31_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c31_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 3))
except Exception as e:
    print("$name parse:", e); print(open("gpurun_out/c31_$name.err").read()[-500:])
PY
}
b v3_fused A=1
b v3_unfused RAVE_FUSE_TAP_STACK=0
