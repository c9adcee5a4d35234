#!/bin/bash
# MRD static prep + gx_full + stacked fm terms + reparam kernel: GPU suite, default bench, v3 quick bench + trace
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c24_tests.log 2>&1; echo "rc=$?"; tail -4 $O/c24_tests.log | cut -c1-300
echo "== bench"; timeout 400 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c24_bench.json 2> $O/c24_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c24_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()})
except Exception as e:
    print("bench parse:", e); print(open("gpurun_out/c24_bench.err").read()[-600:])
PY
echo "== v3"; timeout 300 python bench.py --config v3 --batch 16 --steps 8 --warmup 3 --quick > $O/c24_bench_v3.json 2> $O/c24_bench_v3.err; echo "exit $?"; tail -c 300 $O/c24_bench_v3.json; tail -3 $O/c24_bench_v3.err
timeout 200 python scripts/trace_step_config.py v3 16 > $O/c24_trace_v3.txt 2>&1; head -4 $O/c24_trace_v3.txt | tail -2
