"""One conv_tc launch shape, three launches, for `ncu --set full --import-source on` + `--page source` (per-instruction
stall samples).  Usage: python scripts/ncu_one.py c1|bwd2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rave_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "c1"
torch.manual_seed(0)
if which == "c1":
    B, Cin, Cout, L, K = 64, 64, 384, 4096, 1
else:
    B, Cin, Cout, L, K = 320, 192, 384, 820, 2
x = torch.randn(B, L, Cin, device="cuda").bfloat16()
wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.05).bfloat16()
pad = 0 if K == 1 else 1
Lout = L + 2 * pad - (K - 1)
oa = torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16)
bwd = which != "c1"
dact = torch.randn(B, Lout, Cout, device="cuda").bfloat16() if bwd else None
fmd = torch.tensor([0.3, -0.2], device="cuda") if bwd else None
bias = None if bwd else torch.randn(Cout, device="cuda")
for _ in range(3):
    ops.conv1d_tc(x, wt, bias, None, 1, 1, (pad, pad), 0 if bwd else 1, 0.2, want_f32=False, want_act=False, out_act=oa,
                  Lout=Lout, dact_src=dact, fm_d=fmd)
torch.cuda.synchronize()
