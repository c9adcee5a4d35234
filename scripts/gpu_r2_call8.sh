#!/bin/bash
# round-2 GPU call 8: full GPU suite (one process per group), bench, step trace
mkdir -p gpurun_out
O=gpurun_out
bash scripts/gpu_tests_isolated.sh > $O/c8_tests.txt 2>&1; cat $O/pytest_iso.txt | grep -E "^==|FAILED|ERROR" | cut -c1-200
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $O/c8_bench.json 2> $O/c8_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c8_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], d["roofline"]["kernel"], d["roofline"].get("frac"), {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()}, d["stock_cudnn_tf32"], d["e2e"])
except Exception as e:
    print("bench parse:", e)
PY
echo "== trace"; timeout 600 python scripts/trace_step.py > $O/c8_trace_step.txt 2>&1; grep "====" $O/c8_trace_step.txt
rm -f $O/pytest_*.log.tmp; du -sh gpurun_out
