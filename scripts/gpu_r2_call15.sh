#!/bin/bash
# round-2 GPU call 15: MRD on the engine: tests, v3 single-GPU bench + trace
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c15_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c15_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c15_$name.log | cut -c1-300; }
t descript tests/test_gpu_descript.py
b() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-cudnn-baseline "$@" > $O/c15_bench_$name.json 2> $O/c15_bench_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c15_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, d["config"].get("tcgen05_engine"), d["config"].get("launch"))
except Exception as e:
    print("$name bench parse:", e); print(open("gpurun_out/c15_bench_$name.err").read()[-1500:])
PY
}
b v3 --config v3 --batch 16 --steps 4 --warmup 3
sed -n '/^echo "== v3 trace"/,/^grep "===="/p' scripts/gpu_r2_call14.sh | sed 's/c14_/c15_/g' > /tmp/v3trace.sh; bash /tmp/v3trace.sh
du -sh gpurun_out
