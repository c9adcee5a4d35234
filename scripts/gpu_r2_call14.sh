#!/bin/bash
# round-2 GPU call 14: v3 on the engine (Snake chains, Descript MPD chain): tests + single-GPU bench lines of configs 4 / 5
mkdir -p gpurun_out
O=gpurun_out
t() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu --no-header -rf -s 2>&1 | tail -${TAILN:-40} > $O/c14_$name.log; echo "== $name: $(grep -E 'passed|failed|error' $O/c14_$name.log | tail -1)"; grep -E "^(FAILED|ERROR)" $O/c14_$name.log | cut -c1-300; }
t engine tests/test_gpu_engine.py
t descript tests/test_gpu_descript.py
t parity_v3 tests/test_gpu_parity.py -k "v3 or v1 or snake"
b() { name=$1; shift; timeout 900 python bench.py --no-cpu-baseline --no-cudnn-baseline "$@" > $O/c14_bench_$name.json 2> $O/c14_bench_$name.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/c14_bench_$name.json").read().strip().splitlines()[-1])
    print("$name", {k: d[k] for k in ("value", "ms_per_step")}, d["config"].get("tcgen05_engine"), d["config"].get("launch"))
except Exception as e:
    print("$name bench parse:", e); print(open("gpurun_out/c14_bench_$name.err").read()[-1500:])
PY
}
b v3 --config v3 --batch 16 --steps 4 --warmup 3
b discrete --config discrete --batch 32 --steps 8 --warmup 4
echo "== v3 trace"; timeout 600 python - > $O/c14_trace_v3.txt 2>&1 <<'PY'
import sys, os, re, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
import rave_b200
from rave_b200 import configs
rave_b200.set_precision("bf16")
torch.manual_seed(0)
model = configs.build_rave("v3", sampling_rate=48000).cuda().train()
model.warmed_up = True
x = torch.randn(16, 1, 65536, device="cuda") * 0.1
for i in range(3):
    model.training_step(x, i)
torch.cuda.synchronize()
for tag, idx in (("G-step", 1), ("D-step", 0)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.training_step(x, idx)
        torch.cuda.synchronize()
    ks = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s_, e_, n in ks:
        n = re.sub(r"\(.*", "", n)[:80]
        agg[n][0] += 1; agg[n][1] += e_ - s_
    print(f"==== {tag}: {len(ks)} kernels, span {(ks[-1][1]-ks[0][0])/1e3:.2f} ms, busy {sum(e-s for s,e,_ in ks)/1e3:.2f} ms")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"   {t/1e3:9.3f} ms x {c:4d}  {n}")
PY
grep "====" $O/c14_trace_v3.txt
du -sh gpurun_out
