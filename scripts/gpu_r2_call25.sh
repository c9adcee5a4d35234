#!/bin/bash
# smem weight-norm backward + static encoder prep: GPU suite, default bench (with / without the smem kernel), v3 + discrete quick
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c25_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c25_tests.log | cut -c1-300
b() { name=$1; shift; env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 > $O/c25_$name.json 2> $O/c25_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c25_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d["ms_per_step"], 3), d["config"].get("tcgen05_engine"))
except Exception as e:
    print("$name parse:", e); print(open("gpurun_out/c25_$name.err").read()[-400:])
PY
}
b v2_smem A=1
b v2_gmem RAVE_WN_SMEM=0
b v2_smem2 A=1
timeout 300 python bench.py --config v3 --batch 16 --steps 8 --warmup 3 --quick > $O/c25_bench_v3.json 2> $O/c25_bench_v3.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c25_bench_v3.json").read().strip().splitlines()[-1]); print("v3", round(d["ms_per_step"], 3), d["config"].get("tcgen05_engine"))
except Exception as e:
    print("v3 parse", e)
PY
timeout 300 python bench.py --config discrete --batch 32 --steps 8 --warmup 3 --quick > $O/c25_bench_discrete.json 2> $O/c25_bench_discrete.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c25_bench_discrete.json").read().strip().splitlines()[-1]); print("discrete", round(d["ms_per_step"], 3))
except Exception as e:
    print("discrete parse", e)
PY
