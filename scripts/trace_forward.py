"""In-graph timeline of the north-star forward (PQMF + encoder + generator, v2, B = 32 x 65536): per-kernel durations from
CUPTI during a CUDA-graph replay, in launch order.  Usage: python scripts/trace_forward.py [bf16|bf16x3] > gpurun_out/trace_fwd.txt"""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import rave_b200
from rave_b200 import configs

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
rave_b200.set_precision(mode)
torch.manual_seed(0)
model = configs.build_rave("v2", sampling_rate=48000).cuda().train()
x = torch.randn(32, 1, 65536, device="cuda") * 0.1
with torch.no_grad():
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            model(x)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        y = model(x)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        g.replay()
        torch.cuda.synchronize()
evs = sorted((e.time_range.start, e.time_range.end, e.name) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA)
t0 = evs[0][0]
print(f"{len(evs)} kernels, span {(evs[-1][1]-t0)/1e3:.3f} ms, busy {sum(e-s for s,e,_ in evs)/1e3:.3f} ms")
last = t0
for s_, e_, n in evs:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("rave::tc::", "").replace("rave::", "")[:64]
    print(f"{(s_-t0):9.1f} us  +gap {(s_-last):5.1f}  dur {(e_-s_):7.1f} us  {n}")
    last = e_
