"""The launch bench.py's `roofline` block names (largest in-step share: conv_tc2_kernel<192,64> on the MSD scale-0 third layer,
192 -> 384 channels, k15 s4, bias + LeakyReLU operand out), three launches over rotating buffers, for
  ncu --set full --clock-control none -k regex:conv_tc2 -c 3 -o ... python scripts/ncu_dominant.py
`scripts/ncu_traffic.py` turns the raw page into profiles/r2_ncu_traffic.json, which bench.py reads for `roofline.traffic`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rave_b200 import ops
torch.manual_seed(0)
B, Cin, Cout, Lin, K, st, pad = 64, 192, 384, 4096, 15, 4, 7
Lout = (Lin + 2 * pad - K) // st + 1
xs = [torch.randn(B, Lin, Cin, device="cuda").bfloat16() for _ in range(3)]
wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.02).bfloat16()
bias = torch.randn(Cout, device="cuda")
oa = [torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
for i in range(3):
    ops.conv1d_tc(xs[i], wt, bias, None, st, 1, (pad, pad), 1, 0.2, want_f32=False, want_act=False, out_act=oa[i], Lout=Lout)
torch.cuda.synchronize()
print("launched", B, Cin, Cout, Lin, Lout, K, st)
