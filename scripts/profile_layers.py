"""Time every library launch of one eager G-step and one eager D-step ALONE (sync + CUDA events around each C-ABI
call: rave_b200._lib.PROFILE), group by entry point + shape, and compare the tcgen05 launches with their roofline.
Usage (GPU box): python scripts/profile_layers.py [B]  > gpurun_out/layers.txt"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rave_b200
from rave_b200 import _lib, configs

PEAK_F, PEAK_B = 1.689e15, 7.0e12
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rave_b200.set_precision("bf16")
torch.manual_seed(0)
model = configs.build_rave("v2", sampling_rate=48000).cuda().train()
model.warmed_up = True
x = torch.randn(B, 1, 65536, device="cuda") * 0.1
for i in range(8):
    model.training_step(x, i)
torch.cuda.synchronize()


def summarize(tag, log):
    agg = collections.OrderedDict()
    for name, ints, ptrs, ms in log:
        k = (name, ints, ptrs)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"==== {tag}: {len(log)} library launches, {tot:.3f} ms (each timed alone, warm L2)")
    fam = collections.defaultdict(float)
    rows = []
    for (name, ints, ptrs), (c, ms) in agg.items():
        fam[name] += ms
        extra = ""
        t = 0.0
        if name == "rave_conv1d_tc_fwd":
            Bc, Cin, Lin, pitch, Cout, Lout, K, stride, dil, pad = ints[:10]
            fl = 2.0 * Bc * Lout * Cout * Cin * K
            by = 2.0 * Bc * Lin * Cin / max(1, 1) + Bc * Lout * Cout * 2.0
            # pointer string: xa wt bias res res_bf16 dact res_act out_f32 out_act
            if len(ptrs) >= 9:
                if ptrs[7] == "P": by += Bc * Lout * Cout * 4.0
                if ptrs[3] == "P": by += Bc * Lout * Cout * 4.0
                if ptrs[4] == "P": by += Bc * Lout * Cout * 2.0
                if ptrs[5] == "P": by += Bc * Lout * Cout * 2.0
                if ptrs[6] == "P": by += Bc * Lout * Cout * 2.0
            t = max(fl / PEAK_F, by / PEAK_B) * 1e3
            extra = f" GF={fl/1e9:7.1f} MB={by/1e6:7.1f} roof={t*1e3:7.1f}us frac={t*c/ms:5.2f} TF={fl*c/ms/1e9:6.0f}"
        elif name == "rave_conv1d_tc_wgrad":
            Bc, Cm, Lp, pp, Cn, Lq, qp, K, stride, dil, pad = ints[:11]
            fl = 2.0 * Bc * Lp * Cm * Cn * K
            by = 2.0 * Bc * (Lp * Cm + Lq * Cn)
            t = max(fl / PEAK_F, by / PEAK_B) * 1e3
            extra = f" GF={fl/1e9:7.1f} MB={by/1e6:7.1f} roof={t*1e3:7.1f}us frac={t*c/ms:5.2f} TF={fl*c/ms/1e9:6.0f}"
        rows.append((ms, c, name, ints, ptrs, extra, t * c))
    for n, ms in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"  {ms:8.3f} ms  {ms/tot*100:5.1f}%  {n}")
    tot_roof = sum(r[6] for r in rows)
    tot_tc = sum(r[0] for r in rows if r[6] > 0)
    print(f"  tcgen05 launches: {tot_tc:.3f} ms measured vs {tot_roof:.3f} ms roofline "
          f"(max(flops/{PEAK_F/1e12:.0f} TF, bytes/{PEAK_B/1e12:.1f} TB/s) per launch)")
    print("  -- per shape, sorted by time above the roofline")
    for ms, c, name, ints, ptrs, extra, roof in sorted(rows, key=lambda r: -(r[0] - r[6]))[:25]:
        print(f"  +{(ms-roof)*1e3:7.1f}us  {ms*1e3:8.1f}us x{c:3d} {name[5:]:20s} {ints} {ptrs}{extra}")
    print("  -- per shape (sorted by total time)")
    for ms, c, name, ints, ptrs, extra, roof in sorted(rows, key=lambda r: -r[0])[:70]:
        print(f"  {ms*1e3:8.1f}us x{c:3d} avg {ms/c*1e3:7.1f}us {name[5:]:24s} {ints} {ptrs}{extra}")


for tag, step_idx in (("G-step", 1), ("D-step", 0)):
    _lib.PROFILE = []
    model.training_step(x, step_idx)
    torch.cuda.synchronize()
    log, _lib.PROFILE = _lib.PROFILE, None
    summarize(tag, log)
