#!/bin/bash
# round-2 GPU call 5: L2-prefetched chunk loads; graph-timed layer table + ablations; reworked golden-step test
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-60} > $O/c5_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c5_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c5_$name.log | cut -c1-200; }
TAILN=40 t tc tests/test_gpu_tc.py
TAILN=60 t engine tests/test_gpu_engine.py
TAILN=120 t step tests/test_gpu_parity.py -k "training_step_matches_reference_goldens"
TAILN=80 t disc tests/test_gpu_discrete.py
grep -E "rel-L2|worst|cos" $O/c5_step.log | head
echo "== layer table (graph-timed)"; REPS=10 timeout 600 python scripts/ncu_layers.py 2>&1 | tee $O/c5_layers.txt
for f in "c1 as 64" "J=2, mask" "384->4x192" "msd_192_384"; do timeout 300 python scripts/ablate_tc.py "$f" 2>&1 | tee -a $O/c5_ablate.txt; done
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c5_bench.json 2> $O/c5_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c5_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"]["frac"], {k: v for k, v in d["forward_pqmf_enc_gen"]["modes"].items()})
except Exception as e:
    print("bench parse:", e)
PY
du -sh gpurun_out
