#!/bin/bash
# round-2 GPU call 3: TMA-staged epilogue (tests, layer table, ncu), previously failing tests, step trace, bench
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-60} > $O/c3_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c3_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c3_$name.log | cut -c1-200; }
TAILN=40 t tc_etma tests/test_gpu_tc.py
TAILN=60 t engine_etma tests/test_gpu_engine.py
t fail_step tests/test_gpu_parity.py -k "training_step_matches_reference_goldens"
t fail_x3 "tests/test_gpu_x3.py::test_forward_x3_within_north_star_tolerance"
t fail_disc tests/test_gpu_discrete.py -k "training_steps_run_in_bf16_and_graphs"
t fail_mrd tests/test_gpu_descript.py -k "mrd_vs_oracle"
echo "== layer table, TMA-staged epilogue"; REPS=20 timeout 300 python scripts/ncu_layers.py 2>&1 | tee $O/c3_layers_etma1.txt | head -8
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c3_bench.json 2> $O/c3_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c3_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")}, d["step_roofline"], {k: v for k, v in d["forward_pqmf_enc_gen"].items() if k != "note"})
except Exception as e:
    print("bench parse:", e)
PY
echo "== trace"; timeout 600 python scripts/trace_step.py > $O/c3_trace_step.txt 2>&1; head -3 $O/c3_trace_step.txt
echo "== ncu"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_tc2|dilated_unit|pqmf' -c 60 -o /tmp/r2_ncu_c3 -f python scripts/ncu_layers.py > $O/c3_ncu.log 2>&1
ncu -i /tmp/r2_ncu_c3.ncu-rep --page raw --csv > $O/r2_ncu_c3_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep
du -sh gpurun_out
