#!/bin/bash
# round-2 GPU call 9: forward timeline, source-level profile of the fused unit kernel, engine tests with 1024-thread wn bwd
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c9_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c9_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c9_$name.log | cut -c1-200; }
t engine tests/test_gpu_engine.py
t descript tests/test_gpu_descript.py
echo "== forward trace"; timeout 300 python scripts/trace_forward.py bf16 > $O/c9_trace_fwd_bf16.txt 2>&1; head -2 $O/c9_trace_fwd_bf16.txt | tail -1
timeout 300 python scripts/trace_forward.py bf16x3 > $O/c9_trace_fwd_x3.txt 2>&1; head -2 $O/c9_trace_fwd_x3.txt | tail -1
echo "== ncu unit kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dilated_unit -s 2 -c 1 -o /tmp/unit96 -f python scripts/ncu_layers.py > $O/c9_ncu_unit.log 2>&1
ncu -i /tmp/unit96.ncu-rep --page source --csv > $O/c9_src_unit96.csv 2>/dev/null
ncu -i /tmp/unit96.ncu-rep --page raw --csv > $O/c9_raw_unit96.csv 2>/dev/null
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline --no-cudnn-baseline > $O/c9_bench.json 2> $O/c9_bench.err; echo "exit $?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c9_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step")})
except Exception as e:
    print("bench parse:", e)
PY
du -sh gpurun_out
