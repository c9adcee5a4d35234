#!/bin/bash
# staged gather_c1: GPU suite, micro timings, quick bench A/B, ncu of the small staged kernels
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c27_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c27_tests.log | cut -c1-300
timeout 200 python scripts/time_small.py > $O/c27_small_new.txt 2>&1; grep -E "gather|weight" $O/c27_small_new.txt
b() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 > $O/c27_$name.json 2> $O/c27_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c27_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 3))
except Exception as e:
    print("$name parse:", e); print(open("gpurun_out/c27_$name.err").read()[-400:])
PY
}
b v2_new A=1
b v2_old_c1 RAVE_C1_STAGED=0
b v2_new2 A=1
timeout 240 ncu --set full --clock-control none -k regex:'mt_wn_bwd|c1_staged' -c 14 -o /tmp/c27_small -f python scripts/time_small.py > $O/c27_ncu.log 2>&1
ncu -i /tmp/c27_small.ncu-rep --page raw --csv > $O/r2_ncu_raw_small_staged.csv 2>/dev/null
ls -la /tmp/*.ncu-rep; wc -c $O/r2_ncu_raw_small_staged.csv
