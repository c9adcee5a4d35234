#!/bin/bash
# round-2 GPU call 7: specialised / packed-math TMA-staged epilogue, 4 accumulator stages
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-60} > $O/c7_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c7_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c7_$name.log | cut -c1-200; }
TAILN=40 t tc tests/test_gpu_tc.py
TAILN=60 t engine tests/test_gpu_engine.py
TAILN=60 t step tests/test_gpu_parity.py -k "training_step_matches_reference_goldens"
echo "== layer table (graph-timed)"; REPS=10 timeout 600 python scripts/ncu_layers.py 2>&1 | tee $O/c7_layers.txt | head -9
for f in "c1 as 64" "J=2, mask" "384->4x192"; do timeout 300 python scripts/ablate_tc.py "$f" 2>&1 | grep -E "==|default|per-thread|nothing|no named|no TMEM" | tee -a $O/c7_ablate.txt; done
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c7_bench.json 2> $O/c7_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c7_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], {k: (v.get("ms"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()})
except Exception as e:
    print("bench parse:", e)
PY
du -sh gpurun_out
