"""CUPTI kernel summary of one eager G-step and one eager D-step of a named configuration (bf16 mode).
Usage: python scripts/trace_step_config.py v3 16 > gpurun_out/trace_v3.txt"""
import sys, os, re, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import rave_b200
from rave_b200 import configs

name = sys.argv[1] if len(sys.argv) > 1 else "v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rave_b200.set_precision("bf16")
torch.manual_seed(0)
kw = dict(padding_mode="causal") if name == "discrete" else {}
model = configs.build_rave(name, sampling_rate=48000, **kw).cuda().train()
model.warmed_up = True
x = torch.randn(B, 1, 65536, device="cuda") * 0.1
for i in range(3):
    model.training_step(x, i)
torch.cuda.synchronize()
for tag, idx in (("G-step", 1), ("D-step", 0)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        model.training_step(x, idx)
        torch.cuda.synchronize()
    ks = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
                if e.device_type == torch.autograd.DeviceType.CUDA)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s_, e_, n in ks:
        n = re.sub(r"\(.*", "", n)[:90]
        agg[n][0] += 1
        agg[n][1] += e_ - s_
    print(f"==== {name} {tag}: {len(ks)} kernels, span {(ks[-1][1]-ks[0][0])/1e3:.2f} ms, busy {sum(e-s for s,e,_ in ks)/1e3:.2f} ms")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"   {t/1e3:9.3f} ms x {c:4d}  {n}")
