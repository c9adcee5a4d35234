"""Run with RAVE_TC_2CTA=1: the cta_group::2 variant of the tcgen05 conv against the CPU oracle on
bf16-rounded operands (same check as tests/test_gpu_tc.py), plus a timing of the MSD 384->768 layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import rave_oracle as O
from rave_b200 import ops

assert os.environ.get("RAVE_TC_2CTA", "1") == "1"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


CASES = [
    (2, 64, 64, 512, 3, 1, 1, (1, 1), False, False),
    (2, 192, 192, 1000, 3, 1, 9, (9, 9), True, True),
    (3, 128, 256, 640, 8, 4, 1, (3, 4), False, False),
    (4, 768, 1536, 64, 4, 2, 1, (1, 2), False, False),
    (5, 1536, 256, 32, 3, 1, 1, (1, 1), False, False),
    (3, 384, 768, 1024, 15, 4, 1, (7, 7), True, False),
]
worst = 0.0
for case in CASES:
    B, Cin, Cout, L, K, stride, dil, pad, use_bias, use_res = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, L, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    bias = torch.randn(Cout, generator=g) if use_bias else None
    xa = O.leaky_relu(x, 0.2).bfloat16().float()
    wr = w.bfloat16().float()
    y_exact = O.conv1d(xa, wr, bias, stride, dil, pad)
    res = torch.randn(y_exact.shape, generator=g) if use_res else None
    if use_res:
        y_exact = y_exact + res
    xa_cl, _ = ops.ncl_to_cl(x.cuda(), ops.ACT_LEAKY, 0.2)
    wt = ops.weight_to_tapmajor_bf16(w.cuda())
    res_cl = res.permute(0, 2, 1).contiguous().cuda() if use_res else None
    out_f32, out_act = ops.conv1d_tc(xa_cl, wt, bias.cuda() if use_bias else None, res_cl, stride, dil, pad,
                                     ops.ACT_LEAKY, 0.2, want_f32=True, want_act=True)
    torch.cuda.synchronize()
    y = ops.cl_to_ncl(out_f32)
    r = rel(y, y_exact)
    worst = max(worst, r)
    print(case, "rel", f"{r:.2e}", flush=True)
    assert r < 2e-5, (case, r)

B, Cin, Cout, Lin, K, stride, pad = 64, 384, 768, 1024, 15, 4, 7
Lout = (Lin + 2 * pad - K) // stride + 1
x = torch.randn(B, Lin, Cin, device="cuda").bfloat16()
wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.02).bfloat16()
of = torch.empty(B, Lout, Cout, device="cuda")
oa = torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.conv1d_tc(x, wt, None, None, stride, 1, (pad, pad), 1, 0.2, want_f32=False, want_act=False, out_f32=of,
                  out_act=oa, Lout=Lout)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.conv1d_tc(x, wt, None, None, stride, 1, (pad, pad), 1, 0.2, want_f32=False, want_act=False, out_f32=of,
                  out_act=oa, Lout=Lout)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"2CTA msd_384_768: {ms*1e3:.1f} us  {2.0*B*Lout*Cout*Cin*K/ms/1e9:.0f} TFLOP/s; worst rel {worst:.2e}")
print("2CTA OK")
