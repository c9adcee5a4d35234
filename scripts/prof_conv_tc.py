"""Run a few representative tcgen05 launches in isolation (for `ncu --set full -k regex:conv_tc|wgrad_tc`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rave_b200 import ops

torch.manual_seed(0)
SHAPES = [
    # name, B, Cin, Cout, Lin, K, stride, dil, pad
    ("msd_384_768_k15s4", 64, 384, 768, 1024, 15, 4, 1, 7),
    ("msd_96_192_k15s4", 64, 96, 192, 16384, 15, 4, 1, 7),
    ("unit_c96_k3", 32, 96, 96, 4096, 3, 1, 1, 1),
    ("unit_c768_k3", 32, 768, 768, 64, 3, 1, 1, 1),
]
for name, B, Cin, Cout, Lin, K, stride, dil, pad in SHAPES:
    x = torch.randn(B, Lin, Cin, device="cuda").bfloat16()
    wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.02).bfloat16()
    Lout = (Lin + 2 * pad - dil * (K - 1) - 1) // stride + 1
    of = torch.empty(B, Lout, Cout, device="cuda")
    oa = torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv1d_tc(x, wt, None, None, stride, dil, (pad, pad), 1, 0.2, want_f32=False, want_act=False,
                      out_f32=of, out_act=oa, Lout=Lout)
    g = torch.randn(B, Lout, Cout, device="cuda").bfloat16()
    for _ in range(2):
        ops.conv1d_tc_wgrad(g, x, K, stride, dil, pad)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv1d_tc(x, wt, None, None, stride, dil, (pad, pad), 1, 0.2, want_f32=False, want_act=False,
                      out_f32=of, out_act=oa, Lout=Lout)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * Lout * Cout * Cin * K
    by = 2.0 * B * Lin * Cin + B * Lout * Cout * 6 + 2.0 * K * Cout * Cin
    e0.record()
    for _ in range(10):
        ops.conv1d_tc_wgrad(g, x, K, stride, dil, pad)
    e1.record()
    torch.cuda.synchronize()
    msw = e0.elapsed_time(e1) / 10
    print(f"{name}: fwd {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TFLOP/s  {by/ms/1e6:.0f} GB/s | wgrad {msw*1e3:.1f} us "
          f"{fl/msw/1e9:.0f} TFLOP/s", flush=True)
