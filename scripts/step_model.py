"""Per-layer cost model of the v2 phase-2 training step (B=32x65536): FLOPs, algorithmic HBM bytes and the
roofline lower bound of every tcgen05 launch family.  Runs on CPU (planning only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rave_b200
from rave_b200 import engine, configs

PEAK_F, PEAK_B = 1.689e15, 7.0e12
B, T = 32, 65536

def chain_cost(name, specs, Bc, L0, wgrad=True, dgrad_first=True, mult=1.0):
    L = L0
    rows = []
    for i, s in enumerate(specs):
        Lout = engine._out_len(s, L)
        cin, cout = s.Cin + s.cin_pad, s.Cout + s.cout_pad
        fl = 2.0 * Bc * (Lout if s.kind == 'conv' else L) * cout * cin * s.K
        by_f = 2.0 * Bc * L * cin + Bc * Lout * cout * (2 + (4 if (s.want_f32 or s.is_output) else 0))
        by_d = 2.0 * Bc * Lout * cout + 2.0 * Bc * L * cin * 2      # grad in, mask read, grad out
        by_w = 2.0 * Bc * Lout * cout + 2.0 * Bc * L * cin
        t_f = max(fl / PEAK_F, by_f / PEAK_B)
        t_d = max(fl / PEAK_F, by_d / PEAK_B) if (i > 0 or dgrad_first) else 0
        t_w = max(fl / PEAK_F, by_w / PEAK_B) if wgrad else 0
        rows.append((name, i, s.kind, cin, cout, s.K, s.stride, s.dil, L, Lout, fl, t_f, t_d, t_w))
        L = Lout
    return rows

def main():
    rave_b200.set_precision("bf16")
    model = configs.build_rave("v2", sampling_rate=48000)
    enc = engine.plan_sequential(list(model.encoder.encoder.net)) if hasattr(model.encoder, "encoder") else None
    gen = engine.plan_sequential(list(model.decoder.net))
    allrows = []
    if enc: allrows += chain_cost("enc", enc, B, T // 16)
    allrows += chain_cost("gen", gen, B, T // 2048)
    # discriminators: G-step (fwd + dgrad, no wgrad) every step; D-step (fwd + wgrad + dgrad) every 4th
    disc = model.discriminator
    from rave_b200 import discriminator as D
    nets = []
    for m in disc.modules():
        if isinstance(m, D.ConvNet):
            nets.append(m)
    print("convnets:", len(nets))
    for j, net in enumerate(nets):
        specs = engine.plan_convnet(net.net)
        per = getattr(net, "period", None)
        allrows += chain_cost(f"disc{j}", specs, 2 * B, T, wgrad=True)
    tot = {}
    print(f"{'chain':7s} {'i':>2s} kind {'cin':>5s} {'cout':>5s} {'K':>3s} {'s':>2s} {'d':>2s} {'Lin':>6s} {'Lout':>6s} {'GF':>8s} {'t_f us':>7s} {'t_d us':>7s} {'t_w us':>7s}")
    for r in allrows:
        print(f"{r[0]:7s} {r[1]:2d} {r[2]:5s} {r[3]:5d} {r[4]:5d} {r[5]:3d} {r[6]:2d} {r[7]:2d} {r[8]:6d} {r[9]:6d} {r[10]/1e9:8.1f} {r[11]*1e6:7.1f} {r[12]*1e6:7.1f} {r[13]*1e6:7.1f}")
        k = r[0][:4]
        t = tot.setdefault(k, [0, 0, 0, 0])
        t[0] += r[10]; t[1] += r[11]; t[2] += r[12]; t[3] += r[13]
    for k, t in tot.items():
        print(k, f"GF {t[0]/1e9:.0f}  fwd {t[1]*1e3:.3f} ms  dgrad {t[2]*1e3:.3f} ms  wgrad {t[3]*1e3:.3f} ms")

main()
