#!/bin/bash
# validation of the final tree: smoke, GPU suite (one process, as the driver runs it), small-kernel timings, default bench line
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 600 python -m pytest tests -x -q -m gpu > $O/c28_tests.log 2>&1; echo "rc=$?"; tail -3 $O/c28_tests.log | cut -c1-300
timeout 200 python scripts/time_small.py > $O/c28_small.txt 2>&1; grep -E "gather|weight" $O/c28_small.txt
echo "== bench"; timeout 500 python bench.py > $O/c28_bench.json 2> $O/c28_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c28_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], d["roofline"]["kernel"], d["roofline"].get("frac"), {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()}, d["stock_cudnn_tf32"]["ms_per_step"], d["cpu_baseline"]["value"], d["e2e"]["value"])
except Exception as e:
    print("bench parse:", e); print(open("gpurun_out/c28_bench.err").read()[-600:])
PY
