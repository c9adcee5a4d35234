"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden fixtures generated from the unmodified reference.

Tolerances (relative L2, fp32 parity mode): forward <= 2e-5, gradients <= 1e-4.  The oracle's own
fp32 floor against an fp64 run of the same modules is ~1e-6.
"""
import os

import pytest
import torch
import torch.nn as nn

from oracle import rave_oracle as O
from tests.conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-5
BWD_TOL = 1e-4


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def dev(t):
    return t.cuda() if torch.is_tensor(t) else t


# ------------------------------------------------------------------------------------ PQMF
@pytest.mark.parametrize("mode", ["centered", "causal"])
def test_pqmf_operators_golden(mode):
    from rave_b200 import cc, pqmf
    g = load("pqmf.pt")
    with cc.configure(padding_mode=mode):
        p = pqmf.CachedPQMF(attenuation=100, n_band=16).cuda()
    x, y, xr = g[mode]["x"], g[mode]["y"], g[mode]["xr"]
    y_gpu = p(x.cuda())
    assert y_gpu.shape == y.shape
    assert rel_l2(y_gpu, y) < 2e-6
    xr_gpu = p.inverse(y.cuda())
    assert xr_gpu.shape == xr.shape
    assert rel_l2(xr_gpu, xr) < 2e-6


def test_pqmf_backward_vs_oracle():
    from rave_b200 import pqmf
    p = pqmf.CachedPQMF(attenuation=100, n_band=16).cuda()
    hk = p.hk.cpu()
    x = torch.randn(3, 1, 4096)
    xo = x.clone().requires_grad_(True)
    yo = O.pqmf_analysis(xo, hk)
    gy = torch.randn_like(yo)
    (gx_o,) = torch.autograd.grad(yo, xo, gy)
    xg = x.cuda().requires_grad_(True)
    yg = p(xg)
    (gx,) = torch.autograd.grad(yg, xg, gy.cuda())
    assert rel_l2(gx, gx_o) < 1e-5
    yb = torch.randn(3, 16, 256)
    ybo = yb.clone().requires_grad_(True)
    so = O.pqmf_synthesis(ybo, hk)
    gs = torch.randn_like(so)
    (gyb_o,) = torch.autograd.grad(so, ybo, gs)
    ybg = yb.cuda().requires_grad_(True)
    sg = p.inverse(ybg)
    (gyb,) = torch.autograd.grad(sg, ybg, gs.cuda())
    assert rel_l2(sg, so) < 2e-6
    assert rel_l2(gyb, gyb_o) < 1e-5


def test_pqmf_non_cached_variant_matches_polyphase_reference():
    from rave_b200 import pqmf
    g = load("pqmf.pt")
    p = pqmf.PQMF(attenuation=100, n_band=16).cuda()
    x, y = g["centered"]["x"], g["centered"]["y"]
    assert rel_l2(p(x.cuda()), y) < 2e-6
    assert rel_l2(p.inverse(y.cuda()), g["polyphase_inverse"]) < 2e-6


def test_pqmf_full_size_properties():
    """BASELINE size (32 x 65536): linearity, and the ~1e-3 near-perfect-reconstruction of
    analysis -> synthesis with the 16-sample delay (size-independent properties)."""
    from rave_b200 import pqmf
    p = pqmf.CachedPQMF(attenuation=100, n_band=16).cuda()
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = (0.5 * torch.randn(32, 1, 65536, generator=g)).clamp(-1, 1).cuda()
    x2 = torch.randn(32, 1, 65536, generator=g).cuda()
    y = p(x)
    assert y.shape == (32, 16, 4096)
    lin = p(x + 0.25 * x2) - (y + 0.25 * p(x2))
    assert lin.abs().max() < 5e-5
    xr = p.inverse(y)
    assert xr.shape == x.shape
    r = rel_l2(xr[..., 16 + 1024:-1024], x[..., 1024:-1024 - 16])
    assert 0.9e-3 < r < 1.1e-3
    # one row against the CPU oracle
    yo = O.pqmf_analysis(x[5:6].cpu(), p.hk.cpu())
    assert rel_l2(y[5:6], yo) < 2e-6


# ------------------------------------------------------------------------------ conv family
CONV_CASES = [
    # B, Cin, Cout, L, K, stride, dil, pad, act
    (2, 16, 24, 200, 7, 1, 1, (3, 3), 0),
    (2, 24, 24, 333, 3, 1, 9, (9, 9), 1),
    (3, 24, 48, 256, 8, 4, 1, (3, 4), 1),
    (2, 8, 16, 64, 4, 2, 1, (1, 2), 1),
    (1, 130, 70, 97, 3, 1, 3, (6, 0), 1),     # causal padding, ragged sizes
    (2, 1, 12, 1000, 15, 4, 1, (7, 7), 0),    # discriminator first layer (Cin = 1)
    (2, 40, 1, 50, 1, 1, 1, (0, 0), 1),       # discriminator last layer (Cout = 1)
    (2, 12, 12, 128, 3, 1, 1, (1, 1), 2),     # Snake prologue
    (1, 3, 5, 9, 3, 1, 1, (1, 1), 0),         # tiny
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_fwd_bwd_vs_oracle(case):
    from rave_b200 import ops
    B, Cin, Cout, L, K, stride, dil, pad, act = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, L, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    alpha = 0.5 + torch.rand(Cin, generator=g)
    use_res = stride == 1 and Cin == Cout and pad[0] + pad[1] == dil * (K - 1)

    def ref(x, w, b, alpha):
        h = x
        if act == 1:
            h = O.leaky_relu(h, 0.2)
        elif act == 2:
            h = O.snake(h, alpha.view(-1, 1))
        y = O.conv1d(h, w, b, stride, dil, pad)
        return y + x if use_res else y

    xo, wo, bo, ao = (t.clone().requires_grad_(True) for t in (x, w, b, alpha))
    yo = ref(xo, wo, bo, ao)
    gy = torch.randn(yo.shape, generator=g)
    grads_o = torch.autograd.grad(yo, [xo, wo, bo] + ([ao] if act == 2 else []), gy)

    xg, wg, bg, ag = (t.cuda().requires_grad_(True) for t in (x, w, b, alpha))
    yg = ops.conv1d(xg, wg, bg, xg if use_res else None, stride, dil, pad, act, 0.2,
                    ag if act == 2 else None)
    assert yg.shape == yo.shape
    assert rel_l2(yg, yo) < FWD_TOL
    grads_g = torch.autograd.grad(yg, [xg, wg, bg] + ([ag] if act == 2 else []), gy.cuda())
    for a, b_, name in zip(grads_g, grads_o, ["dx", "dw", "db", "dalpha"]):
        assert rel_l2(a, b_) < BWD_TOL, name


CONVT_CASES = [
    (2, 32, 16, 40, 8, 4, 2, 1),
    (2, 24, 12, 33, 4, 2, 1, 1),
    (1, 130, 60, 17, 4, 2, 1, 0),
    (2, 12, 6, 20, 8, 4, 2, 2),
]


@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose1d_fwd_bwd_vs_oracle(case):
    from rave_b200 import ops
    B, Cin, Cout, L, K, stride, padding, act = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    x = torch.randn(B, Cin, L, generator=g)
    w = torch.randn(Cin, Cout, K, generator=g) / (Cin * K / stride) ** 0.5
    alpha = 0.5 + torch.rand(Cin, generator=g)

    def ref(x, w, alpha):
        h = x
        if act == 1:
            h = O.leaky_relu(h, 0.2)
        elif act == 2:
            h = O.snake(h, alpha.view(-1, 1))
        return O.conv_transpose1d(h, w, None, stride, padding)

    xo, wo, ao = (t.clone().requires_grad_(True) for t in (x, w, alpha))
    yo = ref(xo, wo, ao)
    gy = torch.randn(yo.shape, generator=g)
    grads_o = torch.autograd.grad(yo, [xo, wo] + ([ao] if act == 2 else []), gy)
    xg, wg, ag = (t.cuda().requires_grad_(True) for t in (x, w, alpha))
    yg = ops.conv_transpose1d(xg, wg, None, stride, padding, act, 0.2, ag if act == 2 else None)
    assert yg.shape == yo.shape
    assert rel_l2(yg, yo) < FWD_TOL
    grads_g = torch.autograd.grad(yg, [xg, wg] + ([ag] if act == 2 else []), gy.cuda())
    for a, b_, name in zip(grads_g, grads_o, ["dx", "dw", "dalpha"]):
        assert rel_l2(a, b_) < BWD_TOL, name


def test_weight_norm_fwd_bwd():
    from rave_b200 import ops
    for shape in [(24, 16, 7), (1536, 768, 4), (5, 3, 1), (32, 16, 5, 1)]:
        v = torch.randn(*shape)
        gshape = (shape[0],) + (1,) * (len(shape) - 1)
        g = torch.rand(*gshape) + 0.5
        vo, go = v.clone().requires_grad_(True), g.clone().requires_grad_(True)
        wo = O.weight_norm(vo, go)
        gw = torch.randn_like(wo)
        dvo, dgo = torch.autograd.grad(wo, [vo, go], gw)
        vg, gg = v.cuda().requires_grad_(True), g.cuda().requires_grad_(True)
        wg = ops.weight_norm(vg, gg)
        dvg, dgg = torch.autograd.grad(wg, [vg, gg], gw.cuda())
        assert rel_l2(wg, wo) < 1e-6
        assert rel_l2(dvg, dvo) < 1e-5 and rel_l2(dgg, dgo) < 1e-5


def test_reparametrize_kernel_vs_reference_formula():
    """VariationalEncoder.reparametrize on the device (one library pass, rave_reparam_fwd) against the reference's
    formula (rave/blocks.py:725-737) evaluated on the CPU: sample, KL term and the gradients of both."""
    from rave_b200.blocks import VariationalEncoder
    torch.manual_seed(5)
    ve = VariationalEncoder(lambda n_channels=1: torch.nn.Identity(), beta=0.7)
    for shape in [(2, 32, 100), (3, 256, 32), (1, 6, 5)]:
        z = 3.0 * torch.randn(*shape)
        z[0, shape[1] // 2:, 0] = 25.0                                   # softplus threshold branch
        eps = torch.randn(shape[0], shape[1] // 2, shape[2])
        zo = z.clone().requires_grad_(True)
        mean, scale = zo.chunk(2, 1)
        std = torch.nn.functional.softplus(scale) + 1e-4
        var = std * std
        so = eps * std + mean
        klo = 0.7 * (mean * mean + var - torch.log(var) - 1).sum(1).mean()
        probe = torch.randn_like(so)
        (go,) = torch.autograd.grad((so * probe).sum() + 1.3 * klo, zo)
        zg = z.cuda().requires_grad_(True)
        sg, klg = ve.reparametrize(zg, eps.cuda())
        (gg,) = torch.autograd.grad((sg * probe.cuda()).sum() + 1.3 * klg, zg)
        assert rel_l2(sg, so) < 1e-6 and abs(float(klg) - float(klo)) < 2e-6 * abs(float(klo))
        assert rel_l2(gg, go) < 1e-5


def test_am_tanh_and_snake():
    from rave_b200 import ops
    x = torch.randn(2, 32, 100)
    xo = x.clone().requires_grad_(True)
    a, b = xo.split(16, 1)
    yo = torch.tanh(a * torch.sigmoid(b))
    gy = torch.randn_like(yo)
    (gxo,) = torch.autograd.grad(yo, xo, gy)
    xg = x.cuda().requires_grad_(True)
    yg = ops.am_tanh(xg)
    (gxg,) = torch.autograd.grad(yg, xg, gy.cuda())
    assert rel_l2(yg, yo) < 1e-6 and rel_l2(gxg, gxo) < 1e-5
    alpha = 0.5 + torch.rand(32)
    xo = x.clone().requires_grad_(True)
    ao = alpha.clone().requires_grad_(True)
    so = O.snake(xo, ao.view(-1, 1))
    gs = torch.randn_like(so)
    gxo, gao = torch.autograd.grad(so, [xo, ao], gs)
    xg = x.cuda().requires_grad_(True)
    ag = alpha.cuda().requires_grad_(True)
    sg = ops.activation(xg, ops.ACT_SNAKE, 0.0, ag)
    gxg, gag = torch.autograd.grad(sg, [xg, ag], gs.cuda())
    assert rel_l2(sg, so) < 1e-6 and rel_l2(gxg, gxo) < 1e-5 and rel_l2(gag, gao) < 1e-4
    # LeakyReLU: the flat 16-byte kernels, including a tail and a buffer that is not 16-byte aligned
    for shape, off in [((3, 5, 7), 0), ((2, 32, 100), 0), ((1, 3, 1001), 1)]:
        n = shape[0] * shape[1] * shape[2]
        buf = torch.randn(n + off, device="cuda")
        xl = buf[off:].view(shape).detach().requires_grad_(True)
        yl = ops.activation(xl, ops.ACT_LEAKY, 0.1)
        gl = torch.randn_like(yl)
        (gxl,) = torch.autograd.grad(yl, xl, gl)
        assert torch.equal(yl, torch.nn.functional.leaky_relu(xl.detach(), 0.1))
        assert torch.equal(gxl, gl * torch.where(xl.detach() > 0, 1.0, 0.1))


# ---------------------------------------------------------------------------- model-level goldens
def _build_autoencoder(name):
    from rave_b200 import configs
    g = load(f"autoencoder_{name}.pt")
    kw = {"v2_tiny": {}, "v2_tiny_causal": dict(padding_mode="causal"), "v3_tiny": dict(name="v3"),
          "v2_small_tiny": dict(ratios=[4, 2, 2, 2])}[name]
    kw = dict(kw)
    arch = kw.pop("name", "v2")
    pq, enc, dec = configs.make_autoencoder(arch, capacity=8, latent_size=16, **kw)
    holder = nn.Module()
    holder.pqmf, holder.encoder, holder.decoder = pq, enc, dec
    holder.load_state_dict(g["state_dict"], strict=True)
    return g, holder.cuda().train()


@pytest.mark.parametrize("name", ["v2_tiny", "v2_tiny_causal", "v3_tiny", "v2_small_tiny"])
def test_autoencoder_golden_forward_backward(name):
    """PQMF -> EncoderV2 -> reparametrize(eps) -> GeneratorV2 -> PQMF^-1 against tensors produced
    by the unmodified reference (oracle/make_golden.py), forward and gradients."""
    from rave_b200.model import _pqmf_decode, _pqmf_encode
    g, m = _build_autoencoder(name)
    x = g["x"].cuda().requires_grad_(True)
    x_mb = _pqmf_encode(m.pqmf, x)
    z = m.encoder(x_mb)
    zs, kl = m.encoder.reparametrize(z, g["eps"].cuda())
    y_mb = m.decoder(zs)
    y = _pqmf_decode(m.pqmf, y_mb, batch_size=x.shape[:-2], n_channels=1)
    assert y.shape == x.shape
    assert rel_l2(x_mb, g["x_mb"]) < 2e-6
    assert rel_l2(z, g["z"]) < FWD_TOL
    assert rel_l2(kl, g["kl"]) < FWD_TOL
    assert rel_l2(y_mb, g["y_mb"]) < FWD_TOL
    assert rel_l2(y, g["y"]) < FWD_TOL
    loss = (y * g["probe"].cuda()).sum()
    params = dict(m.encoder.named_parameters(prefix="encoder"))
    params.update(dict(m.decoder.named_parameters(prefix="decoder")))
    names = sorted(g["grad_params"])
    grads = torch.autograd.grad(loss, [x] + [params[n] for n in names])
    assert rel_l2(grads[0], g["grad_x"]) < BWD_TOL
    worst = max(rel_l2(a, g["grad_params"][n]) for a, n in zip(grads[1:], names))
    assert worst < 5e-4, worst


def test_discriminator_v2_golden_forward_backward():
    from rave_b200 import configs
    g = load("discriminator_v2.pt")
    holder = nn.Module()
    holder.discriminator = configs.make_discriminator_v2(capacity=g["capacity"])
    holder.load_state_dict(g["state_dict"], strict=True)
    disc = holder.discriminator.cuda()
    x = g["x"].cuda().requires_grad_(True)
    feats = disc(x)
    assert len(feats) == 8
    for fa, fb in zip(feats, g["features"]):
        assert len(fa) == 5
        for a, b in zip(fa, fb):
            assert a.shape == b.shape
            assert rel_l2(a, b) < FWD_TOL
    fm, ld, la = O.gan_losses(feats, 1, True)      # pure torch arithmetic on device tensors
    assert rel_l2(fm, g["fm"]) < FWD_TOL and rel_l2(ld, g["loss_dis"]) < FWD_TOL
    tot = fm + ld + la
    pp = dict(disc.named_parameters(prefix="discriminator"))
    names = sorted(g["grad_params"])
    grads = torch.autograd.grad(tot, [x] + [pp[n] for n in names])
    assert rel_l2(grads[0], g["grad_x"]) < BWD_TOL
    worst = max(rel_l2(a, g["grad_params"][n]) for a, n in zip(grads[1:], names))
    assert worst < 5e-4, worst


def test_v2_small_config2_vs_oracle():
    """BASELINE config 2 at reduced batch: v2_small (CAPACITY 48, RATIOS [4,2,2,2]) PQMF+encoder+
    generator forward in fp32 vs the CPU oracle, 1e-4 rel-L2 on z and y."""
    from rave_b200 import configs
    from rave_b200.model import _pqmf_decode, _pqmf_encode
    torch.manual_seed(0)
    pq, enc, dec = configs.make_autoencoder("v2_small")
    holder = nn.Module()
    holder.pqmf, holder.encoder, holder.decoder = pq, enc, dec
    sd = {k: v.detach().clone() for k, v in holder.state_dict().items()}
    gen = torch.Generator().manual_seed(1234)
    x = (0.5 * torch.randn(2, 1, 65536, generator=gen)).clamp(-1, 1)
    eps = torch.randn(2, 128, 128, generator=torch.Generator().manual_seed(4321))
    taps = {}
    y_o = O.rave_forward(x, sd, O.v2_small_config(), eps, taps)
    holder.cuda()
    x_mb = _pqmf_encode(pq, x.cuda())
    z = enc(x_mb)
    zs, _ = enc.reparametrize(z, eps.cuda())
    y = _pqmf_decode(pq, dec(zs), batch_size=x.shape[:-2], n_channels=1)
    assert rel_l2(z, taps["z"]) < 1e-4
    assert rel_l2(y, y_o) < 1e-4


def test_training_step_runs_and_updates():
    """One phase-1 G step, one phase-2 D step, one phase-2 G step on a tiny model."""
    from rave_b200 import configs
    torch.manual_seed(0)
    m = configs.build_rave("v2", capacity=8, latent_size=16, disc_capacity=4).cuda().train()
    x = (0.5 * torch.randn(2, 1, 65536, device="cuda")).clamp(-1, 1)   # multiband STFT needs T/16 > 1024
    w0 = m.decoder.net[0].weight_v.detach().clone()
    d0 = m.discriminator.discriminators[1].layers[0].net[0].weight_v.detach().clone()
    m.training_step(x, 1)
    assert not torch.equal(w0, m.decoder.net[0].weight_v)
    assert torch.equal(d0, m.discriminator.discriminators[1].layers[0].net[0].weight_v)
    m.warmed_up = True
    w1 = m.decoder.net[0].weight_v.detach().clone()
    logs = m.training_step(x, 0)                      # D step
    assert torch.equal(w1, m.decoder.net[0].weight_v)
    assert not torch.equal(d0, m.discriminator.discriminators[1].layers[0].net[0].weight_v)
    assert torch.isfinite(logs["loss_dis"])
    logs = m.training_step(x, 1)                      # G step
    assert not torch.equal(w1, m.decoder.net[0].weight_v)
    for k in ("fullband_spectral_distance", "multiband_spectral_distance", "feature_matching",
              "adversarial", "regularization"):
        assert torch.isfinite(logs[k]), k


def test_descript_mpd_vs_oracle():
    """v3 discriminator, MPD branch (77 % of its FLOPs) on the library kernels vs the CPU oracle;
    also the reference's key names (convs.i.0.*, conv_post.*)."""
    from rave_b200.descript_discriminator import MPD, DescriptDiscriminator
    torch.manual_seed(0)
    mpd = MPD(5)
    sd = {k: v.detach().clone() for k, v in mpd.state_dict().items()}
    assert "convs.0.0.weight_g" in sd and "conv_post.weight_v" in sd and "convs.4.0.bias" in sd
    x = torch.randn(2, 1, 2000)
    want = O.descript_mpd(x, sd, "", 5)
    got = mpd.cuda()(x.cuda())
    assert len(got) == 6
    for a, b in zip(got, want):
        assert a.shape == b.shape
        assert rel_l2(a, b) < FWD_TOL
    dd = DescriptDiscriminator()
    y = torch.randn(2, 1, 300)
    assert rel_l2(dd.preprocess(y.cuda()), O.descript_preprocess(y)) < 1e-6


def test_fused_spectral_distance_vs_oracle():
    """core.AudioDistanceV1 with the fused spectral kernels against the reference golden and
    the oracle's autograd."""
    from functools import partial
    from rave_b200 import core
    g = load("audio_distance.pt")
    dist = core.AudioDistanceV1(partial(core.MultiScaleSTFT, scales=[2048, 1024, 512, 256, 128],
                                        sample_rate=48000, magnitude=True), 1e-7).cuda()
    d = dist(g["x"].cuda(), g["y"].cuda())["spectral_distance"]
    assert rel_l2(d, g["distance"]) < 1e-5
    x = g["x"]
    yo = g["y"].clone().requires_grad_(True)
    (go,) = torch.autograd.grad(O.audio_distance_v1(x, yo), yo)
    yg = g["y"].cuda().requires_grad_(True)
    (gg,) = torch.autograd.grad(dist(x.cuda(), yg)["spectral_distance"], yg)
    assert rel_l2(gg, go) < 1e-4


def test_stft_framing_kernels_vs_torch_stft():
    """rave_stft_frames (+ cuFFT) against torch.stft(center=True, reflect) and its autograd, ragged lengths."""
    from rave_b200 import core
    torch.manual_seed(3)
    for (N, T, scales) in [(3, 4096, [2048, 1024, 512, 256, 128]), (2, 1100, [512, 128]), (2, 1101, [256]), (5, 65536, [2048, 128])]:
        m = core.MultiScaleSTFT(scales, 48000, magnitude=True).cuda()
        x = torch.randn(N, T, device="cuda", requires_grad=True)
        got = m.complex_stfts(x)
        for s, y in zip(scales, got):
            xr = x.detach().clone().requires_grad_(True)
            ref = torch.stft(xr, s, hop_length=s // 4, win_length=s, window=getattr(m, f"window_{s}"), center=True,
                             pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
            assert y.shape == ref.shape
            assert rel_l2(torch.view_as_real(y), torch.view_as_real(ref)) < 1e-5
            wgt = torch.randn_like(torch.view_as_real(ref))
            (g_ref,) = torch.autograd.grad((torch.view_as_real(ref) * wgt).sum(), xr)
            (g_got,) = torch.autograd.grad((torch.view_as_real(y) * wgt).sum(), x, retain_graph=True)
            assert rel_l2(g_got, g_ref) < 1e-5


def test_fused_adam_matches_torch_adam():
    """rave_adam_multi (one launch per <= 96 tensors, device-side lr / step) against torch.optim.Adam."""
    from rave_b200.optim import FusedAdam
    torch.manual_seed(5)
    shapes = [(7,), (96, 1, 15), (192, 96, 5), (300, 300), (1,)] + [(33, 3)] * 120      # > 96 tensors, > one chunk
    ref_p = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    our_p = [p.detach().clone().requires_grad_(True) for p in ref_p]
    ref = torch.optim.Adam(ref_p, 1e-3, (.5, .9))
    ours = FusedAdam(our_p, 1e-3, (.5, .9))
    for it in range(4):
        for a, b in zip(ref_p, our_p):
            g = torch.randn_like(a)
            a.grad = g.clone()
            b.grad = g.clone()
        if it == 2:                                   # the schedule writes the lr in place
            ref.param_groups[0]["lr"] = 5e-4
            ours.param_groups[0]["lr"].fill_(5e-4)
        ref.step()
        ours.step()
    torch.cuda.synchronize()
    for a, b in zip(ref_p, our_p):
        assert rel_l2(b, a) < 2e-6
    assert float(ours.param_groups[0]["step"]) == 4.0


# ------------------------------------------------------------------------------------ training step vs the reference's own
def _run_golden_training_steps(precision):
    """Replays tests/golden/training_step_v2_tiny.pt (three steps of the reference's OWN RAVE.training_step:
    phase-1 G, phase-2 D, phase-2 G) through rave_b200.RAVE.training_step; returns per step (logs, gradients of the
    stepped group as left in .grad, post-step state_dict)."""
    import rave_b200
    from rave_b200 import configs
    g = load("training_step_v2_tiny.pt")
    m = configs.build_rave("v2", capacity=g["cfg"]["capacity"], latent_size=g["cfg"]["latent_size"],
                           disc_capacity=g["disc_capacity"], phase_1_duration=1000)
    m.update_discriminator_every = g["update_discriminator_every"]
    m.load_state_dict(g["state_dict"], strict=True)
    m.cuda().train()
    m.set_receptive_field(*g["receptive_field"])
    rave_b200.set_precision(precision)
    out = []
    prev = g["state_dict"]
    try:
        for st in g["steps"]:
            # every step starts from the REFERENCE's pre-step parameters (as tests/test_oracle_golden.py does for the CPU
            # restatement): the phase-2 generator gradient is a sum of sign terms through a freshly initialised
            # discriminator, and the Adam-sized differences a previous step may leave (sign of near-zero gradients)
            # are not part of what this test pins.  The optimiser moments carry over from our own earlier steps.
            m.load_state_dict({k: v for k, v in prev.items()}, strict=True)
            prev = st["state_dict"]
            m.warmed_up = st["warmed_up"]
            logs = m.training_step(st["x"].cuda(), st["batch_idx"], eps=st["eps"].cuda())
            logs = {k: (v.detach().float().cpu() if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in logs.items()}
            grads = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters() if p.grad is not None}
            sd = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
            out.append((logs, grads, sd))
    finally:
        rave_b200.set_precision("fp32")
    return g, out


def test_training_step_matches_reference_goldens_fp32():
    """fp32 kernels: every logged loss <= 1e-4 of the reference's; gradients within the fp32 conditioning of each
    step (tests/test_oracle_golden.py::GRAD_TOL explains the numbers: the reference's own fp32 gradients are 5e-4 /
    2e-6 / 1.1e-2 from an fp64 evaluation); post-step parameters by counting elements whose Adam update differs."""
    from tests.test_oracle_golden import GRAD_TOL, UPD_FRAC
    g, out = _run_golden_training_steps("fp32")
    prev = g["state_dict"]           # the reference's parameters before the step = ours (re-loaded by the runner)
    for st, (logs, grads, sd) in zip(g["steps"], out):
        for k, want in st["logs"].items():
            if k == "beta_factor":
                continue
            assert k in logs, (st["name"], k)
            assert abs(float(logs[k]) - float(want)) <= 1e-4 * max(abs(float(want)), 1e-3), (st["name"], k,
                                                                                              float(logs[k]), float(want))
        keys = sorted(st["grads"])
        cat = lambda d: torch.cat([d[k].reshape(-1) for k in keys])
        assert set(keys) <= set(grads), (st["name"], sorted(set(keys) - set(grads))[:5])
        r = rel_l2(cat(grads), cat(st["grads"]))
        print(f"{st['name']}: gradient rel-L2 vs the reference {r:.3e}")
        worst = sorted(((rel_l2(grads[k], st["grads"][k]), k) for k in keys), reverse=True)[:4]
        print("   worst tensors:", [(k, f"{e:.2e}") for e, k in worst])
        assert r < GRAD_TOL[st["name"]][0], (st["name"], r)
        lr = 1e-4 if (st["warmed_up"] and st["batch_idx"] % g["update_discriminator_every"] == 0) else 1e-3
        n_bad = n_all = 0
        for k, want in st["state_dict"].items():
            if not want.is_floating_point():
                continue
            upd_ref = (want - prev[k]).double()
            upd = (sd[k] - prev[k]).double()
            if upd_ref.abs().max() == 0:
                assert upd.abs().max() == 0, (st["name"], k)
            else:
                n_bad += int(((upd - upd_ref).abs() > 0.05 * lr).sum())
                n_all += upd.numel()
        # phase-2 generator step: the gradient is a sum of sign terms through a freshly initialised discriminator -- the
        # reference's own fp32 gradient is 1.1e-2 from an fp64 evaluation, the CPU restatement 2.0e-2 from the reference,
        # the GPU kernels 4.6e-2 (other summation order: other sign flips); the fraction of Adam updates that move by
        # more than 5 % of lr scales with that distance (measured on B200: 10.6 %)
        frac = 0.15 if st["name"] == "phase2_gen" else UPD_FRAC[st["name"]]
        assert n_all > 0 and n_bad <= frac * n_all, (st["name"], n_bad, n_all)
        prev = st["state_dict"]


def test_training_step_matches_reference_goldens_bf16():
    """Same three steps on the tcgen05 engine (bf16 operands, fp32 accumulate).  Stated tolerance: losses within
    3 % (spectral distances, KL) / 10 % (feature matching, adversarial: sums of sign-like terms over bf16 features) of
    the reference's; gradient direction cos >= 0.9 over all stepped tensors."""
    g, out = _run_golden_training_steps("bf16")
    for st, (logs, grads, sd) in zip(g["steps"], out):
        for k, want in st["logs"].items():
            if k == "beta_factor":
                continue
            tol = 0.10 if k in ("feature_matching", "adversarial", "pred_fake", "pred_real") else 0.03
            assert abs(float(logs[k]) - float(want)) <= tol * max(abs(float(want)), 1e-3), (st["name"], k,
                                                                                           float(logs[k]), float(want))
        keys = sorted(st["grads"])
        a = torch.cat([grads[k].reshape(-1) for k in keys]).double()
        b = torch.cat([st["grads"][k].reshape(-1) for k in keys]).double()
        cos = float((a @ b) / (a.norm() * b.norm()))
        print(f"{st['name']} (bf16): gradient cos {cos:.4f}, rel-L2 {rel_l2(a, b):.3e}")
        assert cos > 0.9, (st["name"], cos)


def test_eager_bf16_training_uses_updated_weights():
    """FusedAdam writes parameters through raw pointers (no autograd version bump): the engine's cache of prepared
    bf16 weights must be dropped by the optimiser step.  Two eager bf16 steps; the forward after them must equal the
    forward of a twin that loaded the same parameters from scratch (cold cache)."""
    import rave_b200
    from rave_b200 import configs, engine
    torch.manual_seed(0)
    m = configs.build_rave("v2", capacity=16, latent_size=16, disc_capacity=8).cuda().train()
    x = (0.5 * torch.randn(2, 1, 65536, device="cuda")).clamp(-1, 1)
    rave_b200.set_precision("bf16")
    try:
        m.training_step(x, 1)
        m.training_step(x, 1)
        with torch.no_grad():
            y_hot = m.decode(m.encode(x)[:, :16])
            engine.invalidate_prepared()
            y_cold = m.decode(m.encode(x)[:, :16])
        assert torch.equal(y_hot, y_cold)
    finally:
        rave_b200.set_precision("fp32")


def test_noise_generator_v2_vs_oracle():
    """a13 / SURVEY 8f.4: NoiseGeneratorV2 (v2_small's filtered-noise branch) -- strided convs on the library kernels and
    the whole mod_sigmoid -> impulse response -> FFT-convolution tail as ONE library kernel (rave_noise_fir_*) -- against
    the oracle restatement of the reference (4 FFTs), forward and gradients, with the uniform noise injected."""
    from rave_b200 import blocks, cc
    torch.manual_seed(0)
    with cc.configure(conv_bias=False):
        ng = blocks.NoiseGeneratorV2(in_size=48, hidden_size=64, data_size=16, ratios=[2, 2, 2], noise_bands=32)
    sd = {"n." + k: v.detach().clone() for k, v in ng.state_dict().items()}
    B, T = 3, 520
    x = torch.randn(B, 48, T)
    noise = torch.rand(B, T // 8, 16, 8) * 2 - 1
    po = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    xo = x.clone().requires_grad_(True)
    y_o = O.noise_generator_v2(xo, po, "n.", (2, 2, 2), 16, 1, noise)
    probe = torch.randn_like(y_o)
    names = sorted(k for k, v in po.items() if v.requires_grad)
    g_o = torch.autograd.grad((y_o * probe).sum(), [xo] + [po[k] for k in names])
    ng.cuda()
    xg = x.cuda().requires_grad_(True)
    y = ng(xg, noise.cuda())
    assert y.shape == y_o.shape
    assert rel_l2(y, y_o) < FWD_TOL
    pg = dict(ng.named_parameters(prefix="n"))
    g = torch.autograd.grad((y * probe.cuda()).sum(), [xg] + [pg[k] for k in names])
    for k, a, b in zip(["x"] + names, g, g_o):
        assert rel_l2(a, b) < 5e-4, (k, rel_l2(a, b))


def test_v1_encoder_generator_match_reference_golden():
    """a12: the v1 blocks -- Encoder (BatchNorm1d, strided convs, grouped output conv) and Generator
    (UpsampleLayer / ResidualStack, waveform x loudness, filtered-noise branch) -- on the library kernels against the
    reference's own modules (tests/golden/autoencoder_v1_tiny.pt): strict state_dict load, forward and gradients."""
    from rave_b200 import blocks, cc
    g = load("autoencoder_v1_tiny.pt")
    ratios = list(g["ratios"])
    with cc.configure(conv_bias=False):
        enc = blocks.Encoder(data_size=16, capacity=g["capacity"], latent_size=g["latent_size"], ratios=ratios, n_out=2,
                             sample_norm=False, repeat_layers=1)
        dec = blocks.Generator(latent_size=g["latent_size"], capacity=g["capacity"], data_size=16, ratios=ratios[::-1],
                               loud_stride=1, use_noise=True)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in g["state_dict"].items() if k.startswith("encoder.")},
                        strict=True)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in g["state_dict"].items() if k.startswith("decoder.")},
                        strict=True)
    enc.cuda().train()
    dec.cuda().train()
    dec.set_warmed_up(True)
    dec.synth.branches[2].__dict__["_noise_override"] = g["noise"].cuda()
    x = g["x"].cuda().requires_grad_(True)
    z = enc(x)
    assert rel_l2(z, g["z"]) < 2e-5
    zin = g["zin"].cuda().requires_grad_(True)
    y = dec(zin)
    assert rel_l2(y, g["y"]) < 2e-5
    pe, pd = dict(enc.named_parameters()), dict(dec.named_parameters())
    ne = sorted(k[len("encoder."):] for k in g["grads"] if k.startswith("encoder."))
    nd = sorted(k[len("decoder."):] for k in g["grads"] if k.startswith("decoder."))
    ge = torch.autograd.grad((z * g["probe_z"].cuda()).sum(), [x] + [pe[k] for k in ne])
    gd = torch.autograd.grad((y * g["probe_y"].cuda()).sum(), [zin] + [pd[k] for k in nd])
    assert rel_l2(ge[0], g["grad_x"]) < 1e-4 and rel_l2(gd[0], g["grad_zin"]) < 1e-4
    worst = max([(rel_l2(a, g["grads"]["encoder." + k]), "encoder." + k) for k, a in zip(ne, ge[1:])] +
                [(rel_l2(a, g["grads"]["decoder." + k]), "decoder." + k) for k, a in zip(nd, gd[1:])])
    print(f"v1 parameter gradients, worst rel-L2 {worst[0]:.2e} ({worst[1]})")
    assert worst[0] < 5e-4, worst


def test_l1_feature_matching_stats_match_torch():
    """core.mean_difference (norm L1, relative or not) on CUDA goes through rave_l1_stats_f32 / rave_l1_grad_f32: same
    value and gradients as the torch arithmetic of rave/core.py:236-252, also on slices that are not 16-byte aligned."""
    from rave_b200 import core
    torch.manual_seed(11)
    for shape, off in [((4, 32, 1000), 0), ((3, 7, 333), 1), ((2, 1, 5), 0)]:
        base_t = torch.randn(shape[0] * shape[1] * shape[2] + off, device="cuda")
        base_v = torch.randn_like(base_t)
        for relative in (True, False):
            t = base_t[off:].view(shape).clone().requires_grad_(True) if off == 0 else \
                base_t[off:].view(shape).detach().requires_grad_(True)
            v = base_v[off:].view(shape).detach().requires_grad_(True)
            got = core.mean_difference(t, v, "L1", relative)
            gt, gv = torch.autograd.grad(got, [t, v])
            t2, v2 = t.detach().cpu().double().requires_grad_(True), v.detach().cpu().double().requires_grad_(True)
            want = (t2 - v2).abs().mean()
            if relative:
                want = want / t2.abs().mean()
            wt, wv = torch.autograd.grad(want, [t2, v2])
            assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
            assert rel_l2(gt, wt) < 1e-5 and rel_l2(gv, wv) < 1e-5


def _stream_vs_offline(build, x, chunk):
    from rave_b200 import cc
    torch.manual_seed(12)
    off = build()
    cc.use_cached_conv(True)
    try:
        on = build()
    finally:
        cc.use_cached_conv(False)
    on.load_state_dict(off.state_dict(), strict=True)
    off.cuda()
    on.cuda()
    with torch.no_grad():
        y_off = off(x)
        y_on = torch.cat([on(c) for c in x.split(chunk, -1)], -1)
    assert y_on.shape == y_off.shape
    return y_on, y_off, on.cumulative_delay


def test_streaming_cached_convs_reproduce_offline():
    """SURVEY 8f.4 / the property the reference's tests/test_residual.py checks for cached_conv: modules built under
    cc.use_cached_conv(True) and fed consecutive chunks reproduce the offline (non-cached) output, delayed by their
    `cumulative_delay`.  Causal padding (what streaming models are trained with, configs/causal.gin): a whole
    encoder / decoder style stack with zero delay; centred padding: each module kind with its own delay."""
    from rave_b200 import blocks, cc

    def stack(mode):
        def build():
            with cc.configure(conv_bias=True, padding_mode=mode):
                return cc.CachedSequential(
                    blocks.normalization(cc.Conv1d(16, 32, 7, padding=cc.get_padding(7))),
                    blocks.Residual(blocks.DilatedUnit(32, 3, 1)),
                    blocks.Residual(blocks.DilatedUnit(32, 3, 3)),
                    nn.LeakyReLU(.2),
                    blocks.normalization(cc.Conv1d(32, 64, 8, stride=4, padding=cc.get_padding(8, 4))),
                    blocks.Residual(blocks.DilatedUnit(64, 3, 9)),
                    nn.LeakyReLU(.2),
                    blocks.normalization(cc.ConvTranspose1d(64, 32, 8, stride=4, padding=2)),
                    blocks.Residual(blocks.DilatedUnit(32, 3, 1)))
        return build

    x = torch.randn(2, 16, 2048, device="cuda")
    # the transposed conv's symmetric crop is a 2-sample delay even with causal convs around it.  The first samples are
    # a start-up transient in ANY cached-conv implementation: offline, the layers after the transposed conv see zero
    # padding where the stream carries the two cropped-away samples (measured on B200: only outputs 0..1 differ, 7e-4).
    y_on, y_off, _ = _stream_vs_offline(stack("causal"), x, 256)
    warm = 16
    assert rel_l2(y_on[..., 2 + warm:], y_off[..., warm:-2]) < 1e-5

    def one(mod):
        def build():
            with cc.configure(conv_bias=True, padding_mode="centered"):
                return mod()
        return build

    cases = [
        (one(lambda: blocks.Residual(blocks.DilatedUnit(16, 3, 3))), 3),
        (one(lambda: blocks.normalization(cc.Conv1d(16, 32, 8, stride=4, padding=cc.get_padding(8, 4)))), 1),
        (one(lambda: blocks.normalization(cc.ConvTranspose1d(16, 8, 8, stride=4, padding=2))), 2),
        (one(lambda: blocks.normalization(cc.Conv1d(16, 16, 7, padding=cc.get_padding(7)))), 3),
    ]
    for build, want_d in cases:
        y_on, y_off, d = _stream_vs_offline(build, x, 128)
        assert d == want_d, (d, want_d)
        # centred padding: offline pads zeros on the left where the stream has its (zero) cache -> identical from 0 on
        # for a single module, up to the right edge the delayed stream has not produced yet
        assert rel_l2(y_on[..., d:], y_off[..., :y_off.shape[-1] - d]) < 1e-5, (want_d, rel_l2(y_on[..., d:], y_off[..., :-d]))


def test_streaming_cached_pqmf_vs_offline_kernels():
    """CachedPQMF under cc.use_cached_conv(True), chunk by chunk through the cached library convs, against the fused
    offline PQMF kernels (golden-pinned above) delayed by the convs' cumulative delays."""
    from rave_b200 import cc
    from rave_b200.pqmf import CachedPQMF
    off = CachedPQMF(100, 16).cuda()
    cc.use_cached_conv(True)
    try:
        on = CachedPQMF(100, 16).cuda()
    finally:
        cc.use_cached_conv(False)
    assert on.streaming and not off.streaming
    x = torch.randn(2, 1, 16 * 1024, device="cuda")
    with torch.no_grad():
        mb_off = off(x)
        mb_on = torch.cat([on(c) for c in x.split(4096, -1)], -1)
        d_f = on.forward_conv.cumulative_delay
        assert mb_on.shape == mb_off.shape
        assert rel_l2(mb_on[..., d_f + 64:], mb_off[..., 64:mb_off.shape[-1] - d_f]) < 1e-5
        y_off = off.inverse(mb_off)
        y_on = torch.cat([on.inverse(c) for c in mb_off.split(256, -1)], -1)
        d_i = on.inverse_conv.cumulative_delay * 16
        assert y_on.shape == y_off.shape
        assert rel_l2(y_on[..., d_i + 1024:], y_off[..., 1024:y_off.shape[-1] - d_i]) < 1e-5
