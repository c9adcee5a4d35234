"""In-situ GPU timeline of the graphed training step (torch.profiler / CUPTI): per-kernel time INSIDE the graph
replay (warm L2, no serialisation), busy time vs span (= launch-latency gaps), tiny-kernel count.
Usage (GPU box): python scripts/trace_step.py > gpurun_out/trace_step.txt"""
import sys, os, re, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import rave_b200
from rave_b200 import configs
from rave_b200.graphs import GraphedTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rave_b200.set_precision("bf16")
torch.manual_seed(0)
model = configs.build_rave("v2", sampling_rate=48000).cuda().train()
model.warmed_up = True
x = torch.randn(B, 1, 65536, device="cuda") * 0.1
tr = GraphedTrainer(model, x)
for i in range(8):
    tr.step(x, i)
torch.cuda.synchronize()


def short(n):
    if n.startswith("void at::") or "at::native" in n or n.startswith("at::"):
        m = re.sub(r"^void ", "", n)
        m = m.replace("at::native::", "").replace("at::", "")
        f = re.findall(r"(\w*(?:Functor|functor|_kernel_cuda|Ops|Op|index\w*|pad\w*|cat\w*|copy\w*)\w*)", m)
        head = re.sub(r"<.*", "", m)[:36]
        return "at:" + head + " " + " ".join(dict.fromkeys(f[:3]))
    return re.sub(r"\(.*", "", n)[:70]


for tag, idx in (("G-step", 1), ("D-step", 0)):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(x, idx)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = []
    for e in evs:
        tr_ = e.time_range
        ks.append((tr_.start, tr_.end, e.name))
    ks.sort()
    if not ks:
        print(tag, "no CUDA events (CUPTI unavailable?)")
        continue
    span = ks[-1][1] - ks[0][0]
    busy = sum(e - s for s, e, _ in ks)
    tiny = [k for k in ks if k[1] - k[0] < 5.0]
    gaps = 0.0
    last_end = ks[0][0]
    for s, e, _ in ks:
        if s > last_end:
            gaps += s - last_end
        last_end = max(last_end, e)
    print(f"==== {tag}: {len(ks)} kernels, span {span/1e3:.3f} ms, sum of kernel time {busy/1e3:.3f} ms, idle gaps "
          f"{gaps/1e3:.3f} ms, kernels < 5 us: {len(tiny)} ({sum(e-s for s,e,_ in tiny)/1e3:.3f} ms)")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in ks:
        a = agg[short(n)]
        a[0] += 1
        a[1] += e - s
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"  {t/1e3:8.3f} ms {t/busy*100:5.1f}%  x{c:4d}  avg {t/c:7.1f} us  {n}")
