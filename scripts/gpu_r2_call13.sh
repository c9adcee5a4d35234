#!/bin/bash
# round-2 GPU call 13: full GPU suite, per-launch roofline table, default bench line (with cuDNN + CPU arms)
mkdir -p gpurun_out
O=gpurun_out
bash scripts/gpu_tests_isolated.sh > $O/c13_tests.txt 2>&1; grep -E "^==|FAILED|ERROR" $O/pytest_iso.txt | cut -c1-200
grep -h "v1 parameter gradients\|MRD gradient" $O/pytest_*.log | cut -c1-200
echo "== per-launch table"; timeout 600 python scripts/profile_layers.py > $O/c13_layers_roofline.txt 2>&1; grep -E "====|tcgen05 launches" $O/c13_layers_roofline.txt
echo "== bench"; timeout 900 python bench.py > $O/c13_bench.json 2> $O/c13_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c13_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], d["roofline"]["kernel"], d["roofline"].get("frac"), {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()}, d["stock_cudnn_tf32"]["ms_per_step"], d["cpu_baseline"], d["e2e"]["value"], d["config"]["tcgen05_engine"])
except Exception as e:
    print("bench parse:", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c13_bench_ref.json 2> $O/c13_bench_ref.err; echo "exit $?"; cut -c1-400 $O/c13_bench_ref.json
du -sh gpurun_out
