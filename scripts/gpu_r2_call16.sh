#!/bin/bash
# round-2 GPU call 16: MRD time-stack kernel + fused L1 feature matching: tests, v3 bench + trace
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c16_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c16_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c16_$name.log | cut -c1-300; }
t descript tests/test_gpu_descript.py
t parity tests/test_gpu_parity.py -k "l1_feature or training_step or v3 or discriminator"
b() { name=$1; shift; timeout 900 python bench.py --quick "$@" > $O/c16_bench_$name.json 2> $O/c16_bench_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c16_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, d["config"].get("tcgen05_engine"), d["config"].get("launch"))
except Exception as e:
    print("$name bench parse:", e); print(open("gpurun_out/c16_bench_$name.err").read()[-1500:])
PY
}
b v3 --config v3 --batch 16 --steps 4 --warmup 3
timeout 600 python scripts/trace_step_config.py v3 16 > $O/c16_trace_v3.txt 2>&1; grep "====" $O/c16_trace_v3.txt
du -sh gpurun_out
