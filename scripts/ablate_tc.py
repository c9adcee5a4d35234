"""Roofline by ablation for the CTA-pair conv kernel: what bounds the mainloop?  Each row re-times the same launch with
one ingredient removed (RAVE_TC_DBG: 1 = no epilogue stores, 2 = no activation loads, 4 = no weight loads) or with a
shallower ring (RAVE_TC_STAGES) / other L2 promotion.  Results with loads removed are garbage by design.
Usage (GPU box): python scripts/ablate_tc.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from rave_b200 import ops

torch.manual_seed(0)
SHAPES = [
    # name, B, Cin, Cout, Lin, K, stride, dil, pad
    ("msd_384_768_k15s4", 64, 384, 768, 1024, 15, 4, 1, 7),
    ("same, stride 1 (L=256)", 64, 384, 768, 256, 15, 1, 1, 7),
    ("msd_192_384_k15s4", 64, 192, 384, 4096, 15, 4, 1, 7),
    ("mpd_384_768_k5s4", 128, 384, 768, 512, 5, 4, 1, 2),
    ("gemm-like k1 1536->1536", 64, 1536, 1536, 256, 1, 1, 1, 0),
    ("msd_96_192_k15s4 (BK=32)", 64, 96, 192, 16384, 15, 4, 1, 7),
    ("mpd_96_192_k5s4 (BK=32)", 128, 96, 192, 8192, 5, 4, 1, 2),
    ("c1 im2col 16->96 k1", 64, 16, 96, 16384, 1, 1, 1, 0),
    ("c1 as 64->384 k1 (4 positions per row)", 64, 64, 384, 4096, 1, 1, 1, 0),
    ("unit 96->96 k3 B32", 32, 96, 96, 4096, 3, 1, 1, 1),
    ("fused dgrad 192->4x96 (J=4)", 64, 192, 384, 4096, 4, 1, 1, 2),
    ("unit 768->768 k3 L=64", 32, 768, 768, 64, 3, 1, 1, 1),
    ("BWD fused dgrad 192->4x96 (J=2, mask + fm partner)", 320, 192, 384, 820, 2, 1, 1, 1),
    ("BWD fused dgrad 192->4x96 (J=4, mask + fm partner)", 64, 192, 384, 4096, 4, 1, 1, 2),
    ("BWD fused dgrad 384->4x192 (J=4, mask + fm partner)", 64, 384, 768, 1024, 4, 1, 1, 2),
]
CONFIGS = [
    ("default", {}),
    ("4 epilogue warps", {"RAVE_TC_EPIWARPS": "4"}),
    ("8 epilogue warps", {"RAVE_TC_EPIWARPS": "8"}),
    ("no epilogue stores", {"RAVE_TC_DBG": "1"}),
    ("no loads at all", {"RAVE_TC_DBG": "6"}),
    ("no loads, no stores", {"RAVE_TC_DBG": "7"}),
    ("no L2 prefetch of epilogue operands", {"RAVE_TC_DBG": "8"}),
    ("L2 promotion none", {"RAVE_TC_L2PROMO": "0"}),
    ("1-CTA kernel", {"RAVE_TC_2CTA": "0"}),
    ("per-thread epilogue (no TMA staging)", {"RAVE_TC_ETMA": "0"}),
    ("etma: 1 chunk buffer", {"RAVE_TC_ESTAGES": "1"}),
    ("etma: no wait on earlier bulk stores", {"RAVE_TC_DBG": "16"}),
    ("etma: no named barriers", {"RAVE_TC_DBG": "32"}),
    ("etma: no barriers, no store wait", {"RAVE_TC_DBG": "48"}),
    ("etma: no TMEM load", {"RAVE_TC_DBG": "64"}),
    ("etma: no staging stores, no TMA store", {"RAVE_TC_DBG": "1"}),
    ("etma: nothing but the loop (1+32+64)", {"RAVE_TC_DBG": "97"}),
]
KEYS = ["RAVE_TC_STAGES", "RAVE_TC_DBG", "RAVE_TC_L2PROMO", "RAVE_TC_EPIWARPS", "RAVE_TC_ETMA", "RAVE_TC_ESTAGES"]

FILTER = sys.argv[1] if len(sys.argv) > 1 else ""
for name, B, Cin, Cout, Lin, K, stride, dil, pad in SHAPES:
    if FILTER and FILTER not in name:
        continue
    x = torch.randn(B, Lin, Cin, device="cuda").bfloat16()
    wt = (torch.randn(K, Cout, Cin, device="cuda") * 0.05).bfloat16()
    Lout = (Lin + 2 * pad - dil * (K - 1) - 1) // stride + 1
    oa = torch.empty(B, Lout, Cout, device="cuda", dtype=torch.bfloat16)
    bwd = name.startswith("BWD")
    dact = torch.randn(B, Lout, Cout, device="cuda").bfloat16() if bwd else None
    fmd = torch.tensor([0.3, -0.2], device="cuda") if bwd else None
    fl = 2.0 * B * Lout * Cout * Cin * K
    print(f"== {name}: {fl/1e9:.1f} GFLOP, rows {B*Lout}", flush=True)
    for cname, env in CONFIGS:
        if cname == "1-CTA kernel":
            continue            # RAVE_TC_2CTA is latched at first use: covered by scripts/check_2cta.py
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)

        def run():
            ops.conv1d_tc(x, wt, None, None, stride, dil, (pad, pad), 0 if bwd else 1, 0.2, want_f32=False,
                          want_act=False, out_f32=None, out_act=oa, Lout=Lout, dact_src=dact, fm_d=fmd)
        from _timing import graph_time_us
        ms = graph_time_us(lambda i: run(), n=10) * 1e-3
        print(f"   {cname:22s} {ms*1e3:8.1f} us  {fl/ms/1e9:6.0f} TFLOP/s", flush=True)
for k in KEYS:
    os.environ.pop(k, None)
