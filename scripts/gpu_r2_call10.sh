#!/bin/bash
# round-2 GPU call 10: multi-tap wgrad (tests, A/B), wn-bwd block size A/B, forward with static prep
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c10_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c10_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c10_$name.log | cut -c1-200; }
t tc tests/test_gpu_tc.py
t engine tests/test_gpu_engine.py
t step tests/test_gpu_parity.py -k "training_step"
echo "== layer table"; REPS=10 timeout 600 python scripts/ncu_layers.py 2>&1 | grep -E "wgrad|mpd0|pqmf" | tee $O/c10_layers.txt
RAVE_WG_MT=0 REPS=10 timeout 600 python scripts/ncu_layers.py 2>&1 | grep -E "wgrad" | tee -a $O/c10_layers.txt
b() { name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c10_bench_$name.json 2> $O/c10_bench_$name.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c10_bench_$name.json"))
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, {k: (v.get("ms"), v.get("ms_with_weight_prep"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()})
except Exception as e:
    print("$name bench parse:", e)
PY
}
b mt1_wn1024 A=1
b mt0_wn1024 RAVE_WG_MT=0
b mt1_wn256 RAVE_WN_THREADS=256
b mt1_wn1024_again A=1
echo "== trace"; timeout 600 python scripts/trace_step.py > $O/c10_trace_step.txt 2>&1; grep "====" $O/c10_trace_step.txt
du -sh gpurun_out
