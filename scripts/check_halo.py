"""Halo variant of the CTA-pair conv kernel (conv_tc3): correctness against the per-tap kernels (RAVE_TC_HALO=0)
for both descriptor conventions (1: plain start-address offset, 2: + matrix-base-offset field) and timing on the
discriminator / DilatedUnit shapes.  Usage (GPU box): python scripts/check_halo.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rave_b200 import ops

torch.manual_seed(0)
SHAPES = [
    # name, B, Cin, Cout, Lin, K, stride, dil, pad
    ("unit_c64_k3", 2, 64, 64, 512, 3, 1, 1, 1),
    ("unit_c192_k3_d9", 2, 192, 192, 1000, 3, 1, 9, 9),
    ("unit_c96_k3_d3 (BK=32)", 4, 96, 96, 4096, 3, 1, 3, 3),
    ("msd_384_768_k15s4", 64, 384, 768, 1024, 15, 4, 1, 7),
    ("msd_192_384_k15s4", 64, 192, 384, 4096, 15, 4, 1, 7),
    ("msd_96_192_k15s4 (BK=32)", 64, 96, 192, 16384, 15, 4, 1, 7),
    ("mpd_384_768_k5s4", 128, 384, 768, 512, 5, 4, 1, 2),
    ("mpd_96_192_k5s4 (BK=32)", 128, 96, 192, 8192, 5, 4, 1, 2),
    ("unit_c96_k3 B32", 32, 96, 96, 4096, 3, 1, 1, 1),
    ("unit_c192_k3 B32", 32, 192, 192, 1024, 3, 1, 1, 1),
    ("unit_c384_k3 B32", 32, 384, 384, 256, 3, 1, 1, 1),
    ("first_k7 16->96", 32, 16, 96, 4096, 7, 1, 1, 3),
]


def run(mode, x, wt, Lout, stride, dil, pad, n=0):
    os.environ["RAVE_TC_HALO"] = str(mode)
    B = x.shape[0]
    of = torch.empty(B, Lout, wt.shape[1], device="cuda")
    oa = torch.empty(B, Lout, wt.shape[1], device="cuda", dtype=torch.bfloat16)
    ops.conv1d_tc(x, wt, None, None, stride, dil, (pad, pad), 1, 0.2, want_f32=False, want_act=False,
                  out_f32=of, out_act=oa, Lout=Lout)
    torch.cuda.synchronize()
    ms = None
    if n:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ops.conv1d_tc(x, wt, None, None, stride, dil, (pad, pad), 1, 0.2, want_f32=False, want_act=False,
                          out_f32=of, out_act=oa, Lout=Lout)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
    return of, ms


for name, B, Cin, Cout, Lin, K, stride, dil, pad in SHAPES:
    x = torch.randn(B, Lin, Cin, device="cuda").bfloat16()
    wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.05).bfloat16()
    Lout = (Lin + 2 * pad - dil * (K - 1) - 1) // stride + 1
    ref, t0 = run(0, x, wt, Lout, stride, dil, pad, 10)
    fl = 2.0 * B * Lout * Cout * Cin * K
    line = f"{name:28s} per-tap {t0*1e3:7.1f} us {fl/t0/1e9:5.0f} TF |"
    for mode in (1, 2):
        try:
            y, t = run(mode, x, wt, Lout, stride, dil, pad, 10)
            err = ((y - ref).norm() / ref.norm()).item()
            line += f" halo{mode}: err {err:.2e} {t*1e3:7.1f} us {fl/t/1e9:5.0f} TF |"
        except Exception as e:
            line += f" halo{mode}: FAILED {str(e)[:80]} |"
    print(line, flush=True)
