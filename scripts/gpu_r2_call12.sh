#!/bin/bash
# round-2 GPU call 12: ws96 unit kernel with TMA-staged skip / outputs
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c12_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c12_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c12_$name.log | cut -c1-200; }
t tc tests/test_gpu_tc.py
t engine tests/test_gpu_engine.py
t step tests/test_gpu_parity.py -k "training_step"
echo "== layer table"; REPS=10 timeout 600 python scripts/ncu_layers.py 2>&1 | grep -E "fused96" | tee $O/c12_layers.txt
b() { name=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c12_bench_$name.json 2> $O/c12_bench_$name.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/c12_bench_$name.json"))
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, {k: (v.get("ms"), v.get("ms_with_weight_prep"), v.get("frac_of_roofline")) for k, v in d["forward_pqmf_enc_gen"]["modes"].items()})
except Exception as e:
    print("$name bench parse:", e); print(open("gpurun_out/c12_bench_$name.err").read()[-600:])
PY
}
b default A=1
du -sh gpurun_out
