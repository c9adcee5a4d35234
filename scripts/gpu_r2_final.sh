#!/bin/bash
# round-2 final validation: smoke, full GPU suite (one process per group), default bench line, configs 4 / 5 on one GPU
mkdir -p gpurun_out
O=gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash scripts/gpu_tests_isolated.sh > $O/final_tests.txt 2>&1; grep -E "^==|FAILED|ERROR" $O/pytest_iso.txt | cut -c1-200
echo "== bench"; timeout 900 python bench.py > $O/final_bench.json 2> $O/final_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/final_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], d["roofline"]["kernel"], d["roofline"].get("frac"), {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()}, d["stock_cudnn_tf32"]["ms_per_step"], d["cpu_baseline"], d["e2e"]["value"])
except Exception as e:
    print("bench parse:", e)
PY
for c in "discrete 32" "v3 16"; do set -- $c; timeout 900 python bench.py --quick --config $1 --batch $2 --steps 8 --warmup 3 > $O/final_bench_$1.json 2> $O/final_bench_$1.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/final_bench_$1.json").read().strip().splitlines()[-1])
    print("$1", {k: d[k] for k in ("value", "ms_per_step")}, d["config"].get("tcgen05_engine"))
except Exception as e:
    print("$1 parse:", e)
PY
done
du -sh gpurun_out
