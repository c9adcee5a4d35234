"""TEST INFRASTRUCTURE -- generates tests/golden/*.pt by EXECUTING THE UNMODIFIED
REFERENCE (loaded from /root/reference by oracle/ref_loader.py) on seeded inputs, and
asserts that oracle/rave_oracle.py reproduces every tensor.  Run in the build container:

    python -m oracle.make_golden

The fixtures are small (tiny CAPACITY) so they can be committed; the architecture code
path is exactly the one the full-size configs take.
"""
import os
import sys
from functools import partial

import torch
import torch.nn as nn

from oracle import rave_oracle as O
from oracle.ref_loader import load_reference, set_padding_mode

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def check(name, got, want, tol):
    r = rel(got, want)
    status = "ok" if r <= tol else "FAIL"
    print(f"  [{status}] {name}: rel-L2 {r:.3e} (tol {tol:g})")
    assert r <= tol, name


def make_input(B, C, T, seed=1234):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (0.5 * torch.randn(B, C, T, generator=g)).clamp(-1, 1)


def build_ref_autoencoder(R, cfg: O.ArchConfig):
    """Instantiate the reference pqmf/encoder/decoder the way configs/v2.gin binds them."""
    blocks = R.blocks
    act = (lambda dim: blocks.Snake(dim)) if cfg.activation == "snake" else (lambda dim: nn.LeakyReLU(.2))
    adain = (lambda dim: blocks.AdaptiveInstanceNormalization(dim)) if cfg.adain else None
    # configs/snake.gin:10-11 rebinds DilatedUnit's activation too
    orig_du = blocks.DilatedUnit
    blocks.DilatedUnit = partial(orig_du, activation=act)
    try:
        enc = blocks.VariationalEncoder(
            partial(blocks.EncoderV2, data_size=cfg.n_band, capacity=cfg.capacity,
                    ratios=cfg.ratios, latent_size=cfg.latent_size, n_out=cfg.n_out,
                    kernel_size=cfg.kernel_size, dilations=cfg.dilations, activation=act,
                    adain=adain, keep_dim=cfg.keep_dim),
            n_channels=cfg.n_channels)
        dec = blocks.GeneratorV2(data_size=cfg.n_band, capacity=cfg.capacity, ratios=cfg.ratios,
                                 latent_size=cfg.generator_latent, kernel_size=cfg.kernel_size,
                                 dilations=cfg.dilations, amplitude_modulation=cfg.amplitude_modulation,
                                 activation=act, adain=adain, keep_dim=cfg.keep_dim,
                                 n_channels=cfg.n_channels)
    finally:
        blocks.DilatedUnit = orig_du
    return enc, dec


def golden_pqmf(R):
    print("PQMF design + operators")
    torch.manual_seed(0)
    out = {}
    for mode in ("centered", "causal"):
        set_padding_mode(mode)
        p = R.pqmf.CachedPQMF(attenuation=100, n_band=16)
        x = make_input(2, 1, 8192)
        y = p(x)
        xr = p.inverse(y)
        h, hk = O.pqmf_design(100, 16)
        check(f"{mode} design h", h, p.h, 0.0)
        check(f"{mode} design hk", hk, p.hk, 0.0)
        check(f"{mode} analysis", O.pqmf_analysis(x, hk, mode), y, 1e-7)
        check(f"{mode} synthesis", O.pqmf_synthesis(y, hk, mode), xr, 1e-7)
        out[mode] = dict(x=x, y=y.detach(), xr=xr.detach())
        if mode == "centered":
            out["h"] = p.h.clone()
            out["hk"] = p.hk.clone()
            out["forward_conv.weight"] = p.forward_conv.weight.detach().clone()
            out["inverse_conv.weight"] = p.inverse_conv.weight.detach().clone()
            # a4: the three formulations agree
            p2 = R.pqmf.PQMF(100, 16, polyphase=True)
            p3 = R.pqmf.PQMF(100, 16, polyphase=False)
            check("polyphase fwd", O.polyphase_forward(x, hk), p2(x), 1e-7)
            check("classic fwd", O.classic_forward(x, hk), p3(x), 1e-7)
            check("polyphase inv", O.polyphase_inverse(y, hk), p2.inverse(y), 1e-7)
            check("classic inv", O.classic_inverse(y, hk), p3.inverse(y), 1e-7)
            out["polyphase_inverse"] = p2.inverse(y).detach()
    set_padding_mode("centered")
    # config 1: 1 x 131072 round trip property (rel-L2 ~ 1.0e-3, inherent to the 100 dB bank)
    x = make_input(1, 1, 131072, seed=7)
    p = R.pqmf.CachedPQMF(attenuation=100, n_band=16)
    xr = p.inverse(p(x))
    rt = rel(xr[..., 16 + 1024:-1024], x[..., 1024:-1024 - 16])
    print(f"  round-trip rel-L2 (interior, 16-sample delay) = {rt:.4e}")
    out["roundtrip_rel_l2"] = rt
    torch.save(out, os.path.join(GOLDEN, "pqmf.pt"))


def golden_autoencoder(R, name, cfg: O.ArchConfig, B=2, T=8192, seed=0):
    print(f"autoencoder {name}")
    set_padding_mode(cfg.pad_mode)
    torch.manual_seed(seed)
    pq = R.pqmf.CachedPQMF(attenuation=100, n_band=cfg.n_band)
    enc, dec = build_ref_autoencoder(R, cfg)
    if cfg.activation == "snake":
        # default alpha == 1 hides alpha handling; perturb deterministically
        g = torch.Generator().manual_seed(99)
        for m in list(enc.modules()) + list(dec.modules()):
            if isinstance(m, R.blocks.Snake):
                m.alpha.data.copy_(0.5 + torch.rand(m.alpha.shape, generator=g))
    enc.train(), dec.train()
    x = make_input(B, cfg.n_channels, T)
    x_mb = R.model._pqmf_encode(pq, x)
    z = enc(x_mb)
    g = torch.Generator().manual_seed(4321)
    eps = torch.randn(z.shape[0], z.shape[1] // 2, z.shape[2], generator=g)
    mean, scale = z.chunk(2, 1)
    std = nn.functional.softplus(scale) + 1e-4
    zs = eps * std + mean
    kl = (mean * mean + std * std - torch.log(std * std) - 1).sum(1).mean()
    y_mb = dec(zs)
    y = R.model._pqmf_decode(pq, y_mb, batch_size=x.shape[:-2], n_channels=cfg.n_channels)

    sd = {"pqmf." + k: v for k, v in pq.state_dict().items()}
    sd.update({"encoder." + k: v for k, v in enc.state_dict().items()})
    sd.update({"decoder." + k: v for k, v in dec.state_dict().items()})
    sd = {k: v.detach().clone() for k, v in sd.items()}

    taps = {}
    y_o = O.rave_forward(x, sd, cfg, eps, taps)
    check("x_mb", taps["x_mb"], x_mb, 1e-7)
    check("z", taps["z"], z, 1e-6)
    check("y_mb", taps["y_mb"], y_mb, 1e-6)
    check("y", y_o, y, 1e-6)
    zs_o, kl_o = O.reparametrize(z.detach(), eps)
    check("kl", kl_o, kl, 1e-6)

    # backward goldens: d(sum(y * probe))/d{x, a few params}
    gp = torch.Generator().manual_seed(777)
    probe = torch.randn(y.shape, generator=gp)
    xg = x.clone().requires_grad_(True)
    x_mb2 = R.model._pqmf_encode(pq, xg)
    z2 = enc(x_mb2)
    mean2, scale2 = z2.chunk(2, 1)
    zs2 = eps * (nn.functional.softplus(scale2) + 1e-4) + mean2
    y2 = R.model._pqmf_decode(pq, dec(zs2), batch_size=x.shape[:-2], n_channels=cfg.n_channels)
    loss = (y2 * probe).sum()
    params = dict(enc.named_parameters(prefix="encoder"))
    params.update(dict(dec.named_parameters(prefix="decoder")))
    names = sorted(params)
    grads = torch.autograd.grad(loss, [xg] + [params[n] for n in names])
    grad_x = grads[0]
    grad_params = {n: g_.detach().clone() for n, g_ in zip(names, grads[1:])}

    fx = dict(cfg=vars(cfg), state_dict=sd, x=x, eps=eps, x_mb=x_mb.detach(), z=z.detach(),
              zs=zs.detach(), kl=kl.detach(), y_mb=y_mb.detach(), y=y.detach(), probe=probe,
              grad_x=grad_x.detach(), grad_params=grad_params, keys=list(sd.keys()))
    torch.save(fx, os.path.join(GOLDEN, f"autoencoder_{name}.pt"))
    set_padding_mode("centered")


def golden_discriminator_v2(R, capacity=4, B=2, T=8192):
    print("v2 discriminator (MPD + MSD)")
    D = R.discriminator
    torch.manual_seed(5)
    norm = R.blocks.normalization
    D.normalization = lambda m, mode="weight_norm": norm(m, mode)
    try:
        periods_net = partial(D.ConvNet, out_size=1, capacity=capacity, n_layers=4, stride=4,
                              conv=nn.Conv2d, kernel_size=(5, 1))
        scales_net = partial(D.ConvNet, out_size=1, capacity=capacity, n_layers=4, stride=4,
                             conv=nn.Conv1d, kernel_size=15)
        disc = D.CombineDiscriminators([
            partial(D.MultiPeriodDiscriminator, periods=[2, 3, 5, 7, 11], convnet=periods_net),
            partial(D.MultiScaleDiscriminator, n_discriminators=3, convnet=scales_net),
        ], n_channels=1)
    finally:
        D.normalization = norm
    # biases default-init small; make them visible
    x = make_input(2 * B, 1, T + 3, seed=11)   # +3: exercises MPD remainder padding
    feats = disc(x)
    sd = {"discriminator." + k: v.detach().clone() for k, v in disc.state_dict().items()}
    feats_o = O.combine_discriminators_v2(x, sd)
    assert len(feats) == len(feats_o) == 8
    for i, (fa, fb) in enumerate(zip(feats_o, feats)):
        for j, (a, b) in enumerate(zip(fa, fb)):
            assert a.shape == b.shape
            check(f"disc {i}.{j} {tuple(b.shape)}", a, b, 1e-6)
    fm, ld, la = O.gan_losses(feats_o, 1, True)
    # reference training_step arithmetic (rave/model.py:348-379)
    real = [[f[:B] for f in s] for s in feats]
    fake = [[f[B:] for f in s] for s in feats]
    fm_r = 0.
    ld_r = 0.
    la_r = 0.
    for sr, sf in zip(real, fake):
        fm_r = fm_r + sum(map(partial(R.core.mean_difference, norm="L1", relative=True),
                              sr[1:], sf[1:])) / len(sr[1:])
        d_, a_ = R.core.hinge_gan(sr[-1], sf[-1])
        ld_r, la_r = ld_r + d_, la_r + a_
    fm_r = fm_r / len(real)
    check("feature matching", fm, fm_r, 1e-6)
    check("loss_dis", ld, ld_r, 1e-6)
    check("loss_adv", la, la_r, 1e-6)
    # input gradient of loss_dis + fm for the backward parity
    xg = x.clone().requires_grad_(True)
    f2 = disc(xg)
    fm2, ld2, la2 = O.gan_losses(f2, 1, True)
    tot = fm2 + ld2 + la2
    pn = sorted(dict(disc.named_parameters()).keys())
    pp = dict(disc.named_parameters())
    grads = torch.autograd.grad(tot, [xg] + [pp[n] for n in pn])
    fx = dict(capacity=capacity, state_dict=sd, x=x,
              features=[[f.detach() for f in s] for s in feats],
              fm=fm_r.detach(), loss_dis=ld_r.detach(), loss_adv=la_r.detach(),
              grad_x=grads[0].detach(),
              grad_params={"discriminator." + n: g.detach() for n, g in zip(pn, grads[1:])},
              keys=list(sd.keys()))
    torch.save(fx, os.path.join(GOLDEN, "discriminator_v2.pt"))


def build_ref_rave(R, cfg: O.ArchConfig, disc_capacity=4, update_discriminator_every=2, phase_1_duration=1000,
                   kind="v2"):
    """The reference's `rave.RAVE` bound the way configs/{v2,v3,discrete}.gin bind it (v2.gin:53-89, v1.gin:95-112,
    snake.gin, adain.gin, descript_discriminator.gin, discrete.gin:13-49), at the capacities given."""
    D = R.discriminator
    blocks, core = R.blocks, R.core
    norm = blocks.normalization
    D.normalization = lambda m, mode="weight_norm": norm(m, mode)
    snake = kind == "v3"
    act = (lambda dim: blocks.Snake(dim)) if snake else (lambda dim: nn.LeakyReLU(.2))
    adain = (lambda dim: blocks.AdaptiveInstanceNormalization(dim)) if kind == "v3" else None
    if kind == "v3":
        disc = R.descript_discriminator.DescriptDiscriminator
    else:
        periods_net = partial(D.ConvNet, out_size=1, capacity=disc_capacity, n_layers=4, stride=4,
                              conv=nn.Conv2d, kernel_size=(5, 1))
        scales_net = partial(D.ConvNet, out_size=1, capacity=disc_capacity, n_layers=4, stride=4,
                             conv=nn.Conv1d, kernel_size=15)
        disc = partial(D.CombineDiscriminators, [
            partial(D.MultiPeriodDiscriminator, periods=[2, 3, 5, 7, 11], convnet=periods_net),
            partial(D.MultiScaleDiscriminator, n_discriminators=3, convnet=scales_net)])
    enc_v2 = partial(blocks.EncoderV2, data_size=cfg.n_band, capacity=cfg.capacity, ratios=cfg.ratios,
                     latent_size=cfg.latent_size, n_out=1 if kind == "discrete" else 2,
                     kernel_size=cfg.kernel_size, dilations=cfg.dilations, activation=act, adain=adain)
    noise_aug = 0
    if kind == "discrete":
        noise_aug = cfg.latent_size
        enc = partial(blocks.DiscreteEncoder, encoder_cls=enc_v2,
                      vq_cls=partial(R.quantization.ResidualVectorQuantization, num_quantizers=16,
                                     dim=cfg.latent_size, codebook_size=1024),
                      num_quantizers=16, noise_augmentation=noise_aug)
    else:
        enc = partial(blocks.VariationalEncoder, enc_v2)
    dec = partial(blocks.GeneratorV2, data_size=cfg.n_band, capacity=cfg.capacity, ratios=cfg.ratios,
                  latent_size=core.get_augmented_latent_size(cfg.latent_size, noise_aug),
                  kernel_size=cfg.kernel_size, dilations=cfg.dilations,
                  amplitude_modulation=True, activation=act, adain=adain)
    stft = partial(core.MultiScaleSTFT, scales=[2048, 1024, 512, 256, 128], sample_rate=48000, magnitude=True)
    dist = partial(core.AudioDistanceV1, multiscale_stft=stft, log_epsilon=1.0 if kind == "discrete" else 1e-7)
    orig_du = blocks.DilatedUnit
    blocks.DilatedUnit = partial(orig_du, activation=act)        # snake.gin:10-11
    try:
        m = R.model.RAVE(latent_size=cfg.latent_size, sampling_rate=48000, encoder=enc, decoder=dec,
                         discriminator=disc, phase_1_duration=phase_1_duration, gan_loss=core.hinge_gan,
                         valid_signal_crop=True,
                         feature_matching_fun=partial(core.mean_difference, norm="L1", relative=True),
                         num_skipped_features=0 if kind == "discrete" else 1, audio_distance=dist,
                         multiband_audio_distance=dist, weights={"feature_matching": 20},
                         pqmf=partial(R.pqmf.CachedPQMF, attenuation=100, n_band=cfg.n_band),
                         update_discriminator_every=update_discriminator_every, n_channels=1)
    finally:
        blocks.DilatedUnit = orig_du
        D.normalization = norm
    return m


def golden_training_step(R, B=2, T=32768):
    """The reference's OWN `RAVE.training_step` (rave/model.py:288-424) executed under the stubs: a phase-1
    generator step, a phase-2 discriminator step and a phase-2 generator step on the same model, in that order.
    Commits, per step: the batch, the reparametrisation noise the reference drew (global RNG re-seeded right before
    the call; `randn_like` is the first draw), every logged scalar and the full post-step state_dict."""
    print("RAVE.training_step (phase-1 G, phase-2 D, phase-2 G)")
    set_padding_mode("centered")
    cfg = O.ArchConfig(capacity=8, latent_size=16)
    torch.manual_seed(0)
    m = build_ref_rave(R, cfg)
    m.train()
    m.receptive_field[0], m.receptive_field[1] = 1024, 512       # what validation_epoch_end would have measured
    opts = m.configure_optimizers()
    gen_opt, dis_opt = opts[0]["optimizer"], opts[1]["optimizer"]
    logs = {}
    m.optimizers = lambda: (gen_opt, dis_opt)
    m.log = lambda k, v: logs.__setitem__(k, v.detach().clone() if torch.is_tensor(v) else torch.tensor(float(v)))
    m.log_dict = lambda d: [m.log(k, v) for k, v in d.items()]
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    Lz = T // cfg.n_band
    for r in cfg.ratios:
        Lz //= r
    steps = []
    for name, warmed, batch_idx, seed in (("phase1_gen", False, 0, 100), ("phase2_dis", True, 0, 101),
                                          ("phase2_gen", True, 1, 102)):
        m.warmed_up = warmed
        x = make_input(B, 1, T, seed=500 + seed)
        torch.manual_seed(seed)
        eps = torch.randn(B, cfg.latent_size, Lz)
        torch.manual_seed(seed)
        logs.clear()
        m.training_step(x.clone(), batch_idx)
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        # gradients the step's optimiser consumed (zero_grad runs BEFORE backward in the reference, so they are still
        # there): discriminator.* after a D-step, encoder.* / decoder.* after a G-step
        dis_step = warmed and batch_idx % m.update_discriminator_every == 0
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()
                 if p.grad is not None and (k.startswith("discriminator.") == dis_step)
                 and not (warmed and k.startswith("encoder.")) and not k.startswith("pqmf.")}
        steps.append(dict(name=name, warmed_up=warmed, batch_idx=batch_idx, x=x, eps=eps,
                          logs={k: v.clone() for k, v in logs.items()}, state_dict=sd, grads=grads))
        print("  ", name, {k: round(float(v), 6) for k, v in logs.items()})
    # pin the restatement: losses of each step from the state the step started in
    prev = sd0
    for st in steps:
        sdp = {k: v for k, v in prev.items()}
        tot, ldis, parts = O.train_step_losses(st["x"], sdp, cfg, st["eps"], warmed_up=st["warmed_up"], beta=1.0,
                                               receptive_field=(1024, 512), return_parts=True)
        for k, v in parts.items():
            check(f"{st['name']} {k}", v, st["logs"][k], 2e-6)
        if st["warmed_up"]:
            check(f"{st['name']} loss_dis", ldis, st["logs"]["loss_dis"], 2e-6)
        prev = st["state_dict"]
    torch.save(dict(cfg=vars(cfg), disc_capacity=4, update_discriminator_every=2, receptive_field=(1024, 512),
                    state_dict=sd0, steps=steps), os.path.join(GOLDEN, "training_step_v2_tiny.pt"))


def golden_v1(R, capacity=8, latent_size=16, B=2, T=4096, ratios=(4, 4, 2)):
    """a12: the reference's v1 Encoder / Generator (rave/blocks.py:322-503) bound like configs/v1.gin (BatchNorm encoder,
    ResidualStack kernel_sizes [3] / dilations [[1,1],[3,1],[5,1]], NoiseGenerator ratios [4,4,4] / 5 bands), forward in
    training mode and gradients; the generator is run warmed-up with the uniform noise draw reproduced by re-seeding."""
    print("v1 encoder / generator")
    set_padding_mode("centered")
    blocks = R.blocks
    orig_rs, orig_ng = blocks.ResidualStack, blocks.NoiseGenerator
    blocks.ResidualStack = partial(orig_rs, kernel_sizes=[3], dilations_list=[list(d) for d in O.V1_DILATIONS])
    blocks.NoiseGenerator = partial(orig_ng, ratios=[4, 4, 4], noise_bands=5)
    try:
        torch.manual_seed(0)
        enc = blocks.Encoder(data_size=16, capacity=capacity, latent_size=latent_size, ratios=list(ratios), n_out=2,
                             sample_norm=False, repeat_layers=1)
        dec = blocks.Generator(latent_size=latent_size, capacity=capacity, data_size=16, ratios=list(ratios)[::-1],
                               loud_stride=1, use_noise=True)
    finally:
        blocks.ResidualStack, blocks.NoiseGenerator = orig_rs, orig_ng
    enc.train()
    dec.train()
    dec.set_warmed_up(True)
    sd = {"encoder." + k: v.detach().clone() for k, v in enc.state_dict().items()}
    sd.update({"decoder." + k: v.detach().clone() for k, v in dec.state_dict().items()})
    x = make_input(B, 16, T // 16, seed=77)
    xg = x.clone().requires_grad_(True)
    z = enc(xg)
    check("v1 encoder", O.encoder_v1(x, sd, "encoder.", ratios), z, 2e-6)
    zin = z[:, :latent_size].detach().clone().requires_grad_(True)
    torch.manual_seed(123)
    y = dec(zin)
    # the reference drew torch.rand_like(ir) inside NoiseGenerator.forward: reproduce the draw for the restatement
    Lz = zin.shape[-1]
    Lh = Lz
    for r in ratios:
        Lh *= r
    torch.manual_seed(123)
    noise = torch.rand(B, Lh // 64, 16, 64) * 2 - 1
    y_o = O.generator_v1(zin.detach(), sd, "decoder.", tuple(ratios)[::-1], warmed_up=True, noise=noise)
    check("v1 generator", y_o, y, 2e-6)
    probe_z, probe_y = torch.randn_like(z), torch.randn_like(y)
    enc_names = [k for k, p in enc.named_parameters()]
    dec_names = [k for k, p in dec.named_parameters()]
    ge = torch.autograd.grad((z * probe_z).sum(), [xg] + [p for _, p in enc.named_parameters()])
    gd = torch.autograd.grad((y * probe_y).sum(), [zin] + [p for _, p in dec.named_parameters()])
    fx = dict(capacity=capacity, latent_size=latent_size, ratios=tuple(ratios), state_dict=sd, x=x, z=z.detach(),
              zin=zin.detach(), y=y.detach(), noise=noise, probe_z=probe_z, probe_y=probe_y,
              grad_x=ge[0].detach(), grad_zin=gd[0].detach(),
              grads={**{"encoder." + k: g.detach() for k, g in zip(enc_names, ge[1:])},
                     **{"decoder." + k: g.detach() for k, g in zip(dec_names, gd[1:])}})
    torch.save(fx, os.path.join(GOLDEN, "autoencoder_v1_tiny.pt"))


def golden_losses(R):
    print("spectral distance (core.AudioDistanceV1)")
    core = R.core
    dist = core.AudioDistanceV1(partial(core.MultiScaleSTFT, scales=[2048, 1024, 512, 256, 128],
                                        sample_rate=48000, magnitude=True), 1e-7)
    x = make_input(2, 1, 16384, seed=21)
    y = make_input(2, 1, 16384, seed=22)
    d = dist(x, y)["spectral_distance"]
    check("audio_distance_v1", O.audio_distance_v1(x, y), d, 1e-6)
    torch.save(dict(x=x, y=y, distance=d), os.path.join(GOLDEN, "audio_distance.pt"))


def golden_state_dict_keys(R):
    """Key lists of the FULL-SIZE configs (names + shapes only): the drop-in contract of
    SURVEY.md App. B.3."""
    print("state_dict key contract (v2 / v3 full size)")
    out = {}
    for name, cfg in (("v2", O.ArchConfig()),
                      ("v2_small", O.v2_small_config()),
                      ("v3", O.ArchConfig(activation="snake", adain=True))):
        torch.manual_seed(0)
        with torch.device("meta"):
            enc, dec = build_ref_autoencoder(R, cfg)
        keys = {"encoder." + k: tuple(v.shape) for k, v in enc.state_dict().items()}
        keys.update({"decoder." + k: tuple(v.shape) for k, v in dec.state_dict().items()})
        out[name] = keys
        print(f"  {name}: {len(keys)} keys")
    # the whole `rave.RAVE` module tree (pqmf.*, encoder.*, decoder.*, discriminator.*, buffers) at full size
    for name, cfg, kind in (("rave_v2", O.ArchConfig(), "v2"),
                            ("rave_v3", O.ArchConfig(activation="snake", adain=True), "v3"),
                            ("rave_discrete", O.ArchConfig(ratios=(4, 4, 2, 2), n_out=1), "discrete")):
        torch.manual_seed(0)
        m = build_ref_rave(R, cfg, disc_capacity=cfg.capacity, kind=kind)
        out[name] = {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
        print(f"  {name}: {len(out[name])} keys, {sum(v.numel() for v in m.state_dict().values()) / 1e6:.2f} M values")
        del m
    torch.save(out, os.path.join(GOLDEN, "state_dict_keys.pt"))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    R = load_reference()
    norm = R.blocks.normalization
    R.blocks.normalization = lambda m, mode="weight_norm": norm(m, mode)  # configs/v1.gin:41
    tiny = dict(capacity=8, latent_size=16)
    golden_pqmf(R)
    golden_autoencoder(R, "v2_tiny", O.ArchConfig(**tiny))
    golden_autoencoder(R, "v2_tiny_causal", O.ArchConfig(pad_mode="causal", **tiny), B=1, T=8192)
    golden_autoencoder(R, "v3_tiny", O.ArchConfig(activation="snake", adain=True, **tiny), B=1, T=8192)
    golden_autoencoder(R, "v2_small_tiny", O.ArchConfig(capacity=8, latent_size=16, ratios=(4, 2, 2, 2)), B=1, T=4096)
    golden_discriminator_v2(R)
    golden_v1(R)
    golden_losses(R)
    golden_training_step(R)
    golden_state_dict_keys(R)
    print("golden fixtures written to", GOLDEN)


if __name__ == "__main__":
    sys.exit(main())
