"""Representative launches of the step in isolation, one of each after a warm-up, for
`ncu --set full --clock-control none -k regex:'conv_tc|wgrad_tc|pqmf|dilated_unit' ...` (scripts/gpu_ncu.sh) and, without ncu, a table
of their times (CUDA events, rotating buffers > L2).  Shapes are the ones of BASELINE config 3 (v2, B = 32 x 65536):
  mpd0_fwd     : MPD period-2 first layer (Cin = 1 read as 4 positions x 16 taps), 235 MB, HBM-bound
  mpd1_dgrad   : fused dgrad of the MPD period-2 second layer (phase-fused, fm gradient + LeakyReLU' in the epilogue)
  msd0_l1_fwd  : MSD scale-0 second layer 96 -> 192 k15 s4 (tensor/HBM balanced)
  unit96_k1    : DilatedUnit 1x1 conv + skip recovered from the operand, C = 96, L = 4096
  unit768_k3   : DilatedUnit k3 conv, C = 768, L = 64 (launch-floor regime)
  pqmf         : analysis + synthesis
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from rave_b200 import ops

torch.manual_seed(0)
dev = "cuda"
REPS = int(os.environ.get("REPS", "1"))


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


def timed(name, fn, byts, flops, nbuf):
    if REPS <= 1:                     # under ncu: plain launches (2 warm-up + 1)
        for i in range(3):
            fn(i % nbuf)
        torch.cuda.synchronize()
        return
    from _timing import graph_time_us
    us = graph_time_us(lambda i: fn(i % nbuf), n=max(REPS, 2))
    ms = us * 1e-3
    print(f"{name:14s} {us:8.1f} us  {byts/ms/1e6:7.0f} GB/s  {flops/ms/1e9:7.0f} TFLOP/s", flush=True)


NB = 3
# ---- mpd0_fwd: [320][820][64] x kron(I4, w[96][16]) -> act bf16 [320][820][384]
B, L = 320, 820
X = [bf(B, L, 64) for _ in range(NB)]
w = bf(1, 384, 64, scale=0.1)
bias = torch.randn(384, device=dev)
oa = [torch.empty(B, L, 384, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
timed("mpd0_fwd", lambda i: ops.conv1d_tc(X[i], w, bias, None, 1, 1, (0, 0), 1, 0.2, want_f32=False, want_act=False,
                                           out_act=oa[i], Lout=L), 2.0 * B * L * (64 + 384), 2.0 * B * L * 64 * 384, NB)
# ---- mpd1_dgrad: g [160][205][192] -> gp [160][205][4*96] with dact + fm gradient (fake half, partner = real half)
Bh, Lq = 160, 205
g = [bf(Bh, Lq + 3, 192) for _ in range(NB)]
wt = bf(2, 384, 192, scale=0.05)
a_full = [bf(2 * Bh, Lq, 384) for _ in range(NB)]
gp = [torch.empty(Bh, Lq, 384, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
fm_d = torch.tensor([1e-3, 2e-3], device=dev)
timed("mpd1_dgrad", lambda i: ops.conv1d_tc(g[i], wt, None, None, 1, 1, (1, 0), 0, 0.2, want_f32=False, want_act=False,
                                             out_act=gp[i], out_rows=Lq, Lout=Lq, Lin=Lq + 1, dact_src=a_full[i][Bh:],
                                             fm_d=fm_d, fm_partner=a_full[i][:Bh]),
      2.0 * Bh * Lq * (192 + 3 * 384), 2.0 * Bh * Lq * 384 * 192 * 2, NB)
# ---- msd0_l1_fwd
B, Cin, Cout, Lin, K, st, pad = 64, 96, 192, 16384, 15, 4, 7
Lout = (Lin + 2 * pad - K) // st + 1
x1 = [bf(B, Lin, Cin) for _ in range(NB)]
w1 = bf(K, Cout, Cin, scale=0.02)
b1 = torch.randn(Cout, device=dev)
o1 = [torch.empty(B, Lout, Cout, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
timed("msd0_l1_fwd", lambda i: ops.conv1d_tc(x1[i], w1, b1, None, st, 1, (pad, pad), 1, 0.2, want_f32=False,
                                              want_act=False, out_act=o1[i], Lout=Lout),
      2.0 * B * (Lin * Cin + Lout * Cout), 2.0 * B * Lout * Cout * Cin * K, NB)
# ---- unit96_k1 / unit96_k3
B, C, L = 32, 96, 4096
xa = [bf(B, L, C) for _ in range(NB)]
xs = [bf(B, L, C) for _ in range(NB)]
wk1 = bf(1, C, C, scale=0.1)
wk3 = bf(3, C, C, scale=0.1)
ou = [torch.empty(B, L, C, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
timed("unit96_k3", lambda i: ops.conv1d_tc(xs[i], wk3, None, None, 1, 3, (3, 3), 1, 0.2, want_f32=False, want_act=False,
                                            out_act=ou[i], Lout=L), 4.0 * B * L * C, 2.0 * B * L * C * C * 3, NB)
timed("unit96_k1", lambda i: ops.conv1d_tc(xa[i], wk1, None, None, 1, 1, (0, 0), 1, 0.2, want_f32=False, want_act=False,
                                            out_act=ou[i], Lout=L, res_act=xs[i], res_slope=0.2),
      6.0 * B * L * C, 2.0 * B * L * C * C, NB)
# ---- unit768_k3 / k1
B, C, L = 32, 768, 64
xd = [bf(B, L, C) for _ in range(NB)]
wd3 = bf(3, C, C, scale=0.03)
wd1 = bf(1, C, C, scale=0.03)
od = [torch.empty(B, L, C, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
timed("unit768_k3", lambda i: ops.conv1d_tc(xd[i], wd3, None, None, 1, 1, (1, 1), 1, 0.2, want_f32=False, want_act=False,
                                             out_act=od[i], Lout=L), 4.0 * B * L * C + 6.0 * C * C, 2.0 * B * L * C * C * 3, NB)
timed("unit768_k1", lambda i: ops.conv1d_tc(xd[i], wd1, None, None, 1, 1, (0, 0), 1, 0.2, want_f32=False, want_act=False,
                                             out_act=od[i], Lout=L, res_act=xd[(i + 1) % NB], res_slope=0.2),
      6.0 * B * L * C + 2.0 * C * C, 2.0 * B * L * C * C, NB)
# ---- fused Residual(DilatedUnit) kernels (csrc/unit_tc.cu) at the three widths, inference form (no a1 write) and training
for (Bu, Cu, Lu, du) in ((32, 96, 4096, 3), (32, 192, 1024, 3), (32, 384, 256, 3)):
    xu = [bf(Bu, Lu, Cu) for _ in range(NB)]
    w3u, w1u = bf(3, Cu, Cu, scale=0.05), bf(1, Cu, Cu, scale=0.05)
    ouu = [torch.empty(Bu, Lu, Cu, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    for keep in (False, True):
        timed(f"fused{Cu}{'_a1' if keep else ''}",
              lambda i: ops.dilated_unit_tc(xu[i], w3u, w1u, du, du, 0.2, 0.2, 1, 0.2, want_a1=keep, out_act=ouu[i]),
              (6.0 if keep else 4.0) * Bu * Lu * Cu + 8.0 * Cu * Cu, 2.0 * Bu * Lu * Cu * Cu * 4, NB)
# ---- split-operand (bf16x3) conv: DilatedUnit k3 at C = 192
x3a = [torch.cat([bf(32, 1024, 192), bf(32, 1024, 192, scale=0.004)], -1).contiguous() for _ in range(NB)]
w3x = torch.cat([bf(3, 192, 192, scale=0.05), bf(3, 192, 192, scale=0.0002)], 0).contiguous()
o3x = [torch.empty(32, 1024, 384, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
timed("x3_unit192_k3", lambda i: ops.conv1d_tc(x3a[i], w3x, None, None, 1, 3, (3, 3), 1, 0.2, want_f32=False,
                                                want_act=False, out_act=o3x[i], Lout=1024, x3=True),
      8.0 * 32 * 1024 * 192, 6.0 * 32 * 1024 * 192 * 192 * 3, NB)
# ---- wgrad of the MSD 96 -> 192 layer (per-tap form)
gw = bf(64, Lout, 192)
timed("wgrad_msd0_l1", lambda i: ops.conv1d_tc_wgrad(gw, x1[i], K, st, 1, pad), 2.0 * 64 * (Lout * 192 + Lin * 96),
      2.0 * 64 * Lout * 192 * 96 * K, NB)
# ---- PQMF
from rave_b200 import cc, pqmf
with cc.configure():
    pq = pqmf.CachedPQMF(attenuation=100, n_band=16).cuda()
xw = [torch.randn(32, 1, 65536, device=dev) for _ in range(NB)]
yb = [torch.randn(32, 16, 4096, device=dev) for _ in range(NB)]
with torch.no_grad():
    timed("pqmf_analysis", lambda i: pq(xw[i]), 8.0 * 32 * 65536, 2.0 * 32 * 65536 * 512, NB)
    timed("pqmf_synthesis", lambda i: pq.inverse(yb[i]), 8.0 * 32 * 65536, 2.0 * 32 * 65536 * 528, NB)
