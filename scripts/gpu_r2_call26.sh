#!/bin/bash
# staged im2col_c1 + unrolled smem weight-norm backward: GPU suite, micro timings of both variants, quick bench A/B
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c26_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c26_tests.log | cut -c1-300
timeout 200 python scripts/time_small.py > $O/c26_small_new.txt 2>&1; cat $O/c26_small_new.txt | tail -12
RAVE_C1_STAGED=0 RAVE_WN_SMEM=0 timeout 200 python scripts/time_small.py > $O/c26_small_old.txt 2>&1; cat $O/c26_small_old.txt | tail -12
b() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 > $O/c26_$name.json 2> $O/c26_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c26_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 3))
except Exception as e:
    print("$name parse:", e); print(open("gpurun_out/c26_$name.err").read()[-400:])
PY
}
b v2_new A=1
b v2_old_c1 RAVE_C1_STAGED=0
b v2_new2 A=1
