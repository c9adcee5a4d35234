"""TEST INFRASTRUCTURE (CPU oracle) -- NOT part of the product.

A functional CPU restatement of the arithmetic of acids-ircam/RAVE's waveform hot
path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file; rave_b200/ never does.

The reference's arithmetic is ATen fp32 on CPU (F.pad + F.conv1d through the
un-vendored `cached-conv>=2.5.0`, requirements.txt:14), so the restatement uses the
same primitive (torch.nn.functional on CPU tensors) and no CUDA.  Every function takes
plain tensors / a state_dict with the REFERENCE's key names, so the product's
state_dict compatibility is exercised by the same tests.

PINNING: oracle/make_golden.py executes the UNMODIFIED reference modules (loaded from
/root/reference under the stubs of oracle/ref_loader.py) and this restatement on the same
seeded inputs/weights, asserts agreement, and writes tests/golden/*.pt.  The reference
itself holds no golden vectors for this path, so the fixtures
generated from the reference ARE the pin.

All functions work in float32 or float64 (the fp64 run is the arbiter between two fp32
paths that disagree at 1e-6).
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------
# cached_conv 2.5.0 non-cached semantics ------------------------
# ----------------------------------------------------------------------------------


def get_padding(kernel_size: int, stride: int = 1, dilation: int = 1,
                mode: str = "centered") -> Tuple[int, int]:
    """cc.get_padding: `stride` is accepted but unused."""
    if kernel_size == 1:
        return (0, 0)
    p = (kernel_size - 1) * dilation + 1
    if mode == "centered":
        return ((p - 1) // 2, p // 2)
    if mode == "causal":
        return (p // 2 + (p - 1) // 2, 0)
    raise ValueError(mode)


def conv1d(x: Tensor, w: Tensor, b: Optional[Tensor] = None, stride: int = 1,
           dilation: int = 1, pad: Tuple[int, int] = (0, 0)) -> Tensor:
    """cc.Conv1d.forward = F.pad(x, pad) then F.conv1d with padding 0."""
    return F.conv1d(F.pad(x, pad), w, b, stride, 0, dilation)


def conv_transpose1d(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int,
                     padding: int) -> Tensor:
    return F.conv_transpose1d(x, w, b, stride, padding)


def weight_norm(v: Tensor, g: Tensor, dim: int = 0) -> Tensor:
    """torch.nn.utils.weight_norm (rave/blocks.py:15-22): w = g * v / ||v|| with the norm
    taken over every dim except `dim` (0 for every conv here, also for ConvTranspose1d
    whose dim 0 is Cin)."""
    dims = [d for d in range(v.dim()) if d != dim]
    return v * (g / v.norm(2, dims, keepdim=True))


def wn_weight(sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """Effective weight of a (possibly weight-normalised) conv stored at `prefix`."""
    if prefix + "weight_v" in sd:
        return weight_norm(sd[prefix + "weight_v"], sd[prefix + "weight_g"], 0)
    return sd[prefix + "weight"]


def leaky_relu(x: Tensor, slope: float = 0.2) -> Tensor:
    return F.leaky_relu(x, slope)


def snake(x: Tensor, alpha: Tensor) -> Tensor:
    """rave/blocks.py:852-860."""
    return x + (alpha + 1e-9).reciprocal() * (alpha * x).sin().pow(2)


# ----------------------------------------------------------------------------------
# PQMF (rave/pqmf.py) --------------------------------------------------------------
# ----------------------------------------------------------------------------------


def kaiser_filter(wc: float, atten: float, N: Optional[int] = None) -> np.ndarray:
    """rave/pqmf.py:55-70 (scipy 1.10 `firwin(..., nyq=pi)` == `fs=2*pi`)."""
    from scipy.signal import firwin, kaiserord
    N_, beta = kaiserord(atten, wc / np.pi)
    N_ = 2 * (N_ // 2) + 1
    N = N if N is not None else N_
    return firwin(N, wc, window=("kaiser", beta), scale=False, fs=2 * np.pi)


def loss_wc(wc, atten, M, N):
    """rave/pqmf.py:73-80."""
    h = kaiser_filter(wc, atten, N)
    g = np.convolve(h, h[::-1], "full")
    g = abs(g[g.shape[-1] // 2::2 * M][1:])
    return np.max(g)


def get_prototype(atten: float, M: int, N: Optional[int] = None) -> np.ndarray:
    """rave/pqmf.py:83-89."""
    from scipy.optimize import fmin
    wc = fmin(lambda w: loss_wc(w, atten, M, N), 1 / M, disp=0)[0]
    return kaiser_filter(wc, atten, N)


def get_qmf_bank(h: Tensor, n_band: int) -> Tensor:
    """rave/pqmf.py:32-52."""
    k = torch.arange(n_band).reshape(-1, 1)
    N = h.shape[-1]
    t = torch.arange(-(N // 2), N // 2 + 1)
    p = (-1) ** k * math.pi / 4
    mod = torch.cos((2 * k + 1) * math.pi / (2 * n_band) * t + p)
    return 2 * h * mod


def center_pad_next_pow_2(x: Tensor) -> Tensor:
    """rave/pqmf.py:20-23."""
    next_2 = 2 ** math.ceil(math.log2(x.shape[-1]))
    pad = next_2 - x.shape[-1]
    return F.pad(x, (pad // 2, pad // 2 + int(pad % 2)))


def pqmf_design(attenuation: float, n_band: int) -> Tuple[Tensor, Tensor]:
    """PQMF.__init__ (rave/pqmf.py:192-207): returns buffers (h[ntaps], hk[n_band, 2^k])."""
    h = torch.from_numpy(get_prototype(attenuation, n_band)).float()
    hk = center_pad_next_pow_2(get_qmf_bank(h, n_band))
    return h, hk


def reverse_half(x: Tensor) -> Tensor:
    """rave/pqmf.py:13-17: odd bands, even time steps are negated."""
    mask = torch.ones_like(x)
    mask[..., 1::2, ::2] = -1
    return x * mask


def make_odd(x: Tensor) -> Tensor:
    """rave/pqmf.py:26-29."""
    if not x.shape[-1] % 2:
        x = F.pad(x, (0, 1))
    return x


def cached_pqmf_weights(hk: Tensor) -> Tuple[Tensor, Tensor]:
    """CachedPQMF.__init__ (rave/pqmf.py:250-273): analysis weight [M,1,K+1], synthesis
    weight hki[m,c,t] = hk[c, K-1-(M t + m)] made odd -> [M,M,K/M+1]."""
    M, K = hk.shape
    hkf = make_odd(hk).unsqueeze(1)
    hki = hk.flip(-1).reshape(M, K // M, M).permute(2, 0, 1)  # "c (t m) -> m c t"
    hki = make_odd(hki)
    return hkf.contiguous(), hki.contiguous()


def pqmf_analysis(x: Tensor, hk: Tensor, mode: str = "centered") -> Tensor:
    """CachedPQMF.forward (rave/pqmf.py:279-283): x[B,1,T] -> [B,M,T/M]."""
    M = hk.shape[0]
    if M == 1:
        return x
    hkf, _ = cached_pqmf_weights(hk.to(x.dtype))
    y = conv1d(x, hkf, None, stride=M, pad=get_padding(hkf.shape[-1], mode=mode))
    return reverse_half(y)


def pqmf_synthesis(y: Tensor, hk: Tensor, mode: str = "centered") -> Tensor:
    """CachedPQMF.inverse (rave/pqmf.py:285-294): y[B,M,L] -> [B,1,L*M]."""
    M = hk.shape[0]
    if M == 1:
        return y
    _, hki = cached_pqmf_weights(hk.to(y.dtype))
    x = reverse_half(y)
    x = conv1d(x, hki, None, pad=get_padding(hki.shape[-1], mode=mode)) * M
    x = x.flip(1)
    x = x.permute(0, 2, 1)
    x = x.reshape(x.shape[0], x.shape[1], -1, M).permute(0, 2, 1, 3)
    return x.reshape(x.shape[0], x.shape[1], -1)


def polyphase_forward(x: Tensor, hk: Tensor) -> Tensor:
    """rave/pqmf.py:92-108 followed by reverse_half (PQMF.forward, 209-222)."""
    M = hk.shape[0]
    B, C, T = x.shape
    xr = x.reshape(B, C, T // M, M).permute(0, 1, 3, 2).reshape(B, C * M, T // M)
    hkr = hk.reshape(M, hk.shape[-1] // M, M).permute(0, 2, 1)  # "c (t m) -> c m t"
    y = F.conv1d(xr, hkr, padding=hkr.shape[-1] // 2)[..., :-1]
    return reverse_half(y)


def classic_forward(x: Tensor, hk: Tensor) -> Tensor:
    """rave/pqmf.py:137-155 followed by reverse_half."""
    y = F.conv1d(x, hk.unsqueeze(1), stride=hk.shape[0], padding=hk.shape[-1] // 2)[..., :-1]
    return reverse_half(y)


def polyphase_inverse(y: Tensor, hk: Tensor) -> Tensor:
    """PQMF.inverse with polyphase=True (rave/pqmf.py:111-134, 224-242)."""
    m = hk.shape[0]
    x = reverse_half(y)
    hkr = hk.flip(-1)
    hkr = hkr.reshape(m, hk.shape[-1] // m, m).permute(2, 0, 1)  # "c (t m) -> m c t"
    pad = hkr.shape[-1] // 2 + 1
    x = F.conv1d(x, hkr, padding=int(pad))[..., :-1] * m
    x = x.flip(1)
    B, CM, T = x.shape
    x = x.reshape(B, CM // m, m, T).permute(0, 1, 3, 2).reshape(B, CM // m, T * m)
    return x[..., 2 * hkr.shape[1]:]


def classic_inverse(y: Tensor, hk: Tensor) -> Tensor:
    """PQMF.inverse with polyphase=False (rave/pqmf.py:158-176)."""
    x = reverse_half(y)
    hkf = hk.flip(-1)
    up = torch.zeros(*x.shape[:2], hk.shape[0] * x.shape[-1]).to(x)
    up[..., ::hk.shape[0]] = x * hk.shape[0]
    return F.conv1d(up, hkf.unsqueeze(0), padding=hkf.shape[-1] // 2)[..., 1:]


def pqmf_encode(x: Tensor, hk: Tensor, mode: str = "centered") -> Tensor:
    """_pqmf_encode (rave/model.py:116-122): [B,n_ch,T] -> [B,n_ch*M,T/M]."""
    batch = x.shape[:-2]
    xm = pqmf_analysis(x.reshape(-1, 1, x.shape[-1]), hk, mode)
    return xm.reshape(*batch, -1, xm.shape[-1])


def pqmf_decode(y: Tensor, hk: Tensor, n_channels: int = 1, mode: str = "centered") -> Tensor:
    """_pqmf_decode (rave/model.py:125-130)."""
    batch = y.shape[:-2]
    x = y.reshape(y.shape[0] * n_channels, -1, y.shape[-1])
    x = pqmf_synthesis(x, hk, mode)
    return x.reshape(*batch, n_channels, -1)


# ----------------------------------------------------------------------------------
# Encoder / generator (rave/blocks.py) ---------------------------------------------
# ----------------------------------------------------------------------------------


class ArchConfig:
    """Effective gin bindings of one configuration."""

    def __init__(self, capacity=96, ratios=(4, 4, 4, 2), latent_size=128, n_out=2,
                 kernel_size=3, dilations=((1, 3, 9), (1, 3, 9), (1, 3, 9), (1, 3)),
                 n_band=16, n_channels=1, activation="leaky", adain=False,
                 amplitude_modulation=True, pad_mode="centered", keep_dim=False,
                 generator_latent=None):
        self.capacity = capacity
        self.ratios = list(ratios)
        self.latent_size = latent_size
        self.n_out = n_out
        self.kernel_size = kernel_size
        if isinstance(dilations[0], int):
            dilations = [list(dilations) for _ in ratios]
        self.dilations = [list(d) for d in dilations]
        self.n_band = n_band
        self.n_channels = n_channels
        self.activation = activation
        self.adain = adain
        self.amplitude_modulation = amplitude_modulation
        self.pad_mode = pad_mode
        self.keep_dim = keep_dim
        self.generator_latent = generator_latent if generator_latent is not None else latent_size


def _act(x: Tensor, sd, prefix: str, cfg: ArchConfig) -> Tensor:
    """The `activation(dim)` module at `prefix`: LeakyReLU(.2) (blocks.py:528) or Snake
    (configs/snake.gin:5-23)."""
    if cfg.activation == "snake":
        return snake(x, sd[prefix + "alpha"])
    return leaky_relu(x, 0.2)


def dilated_unit(x: Tensor, sd, prefix: str, cfg: ArchConfig, dim: int, dilation: int) -> Tensor:
    """Residual(DilatedUnit) (rave/blocks.py:31-45, 83-112); `prefix` ends with
    'aligned.branches.0.net.'."""
    k = cfg.kernel_size
    h = _act(x, sd, prefix + "0.", cfg)
    h = conv1d(h, wn_weight(sd, prefix + "1."), sd.get(prefix + "1.bias"), 1, dilation,
               get_padding(k, dilation=dilation, mode=cfg.pad_mode))
    h = _act(h, sd, prefix + "2.", cfg)
    h = conv1d(h, wn_weight(sd, prefix + "3."), sd.get(prefix + "3.bias"))
    return h + x


def encoder_v2(x: Tensor, sd, prefix: str, cfg: ArchConfig,
               taps: Optional[dict] = None) -> Tensor:
    """EncoderV2.forward (rave/blocks.py:514-596).  `prefix` e.g. 'encoder.encoder.'."""
    p = prefix + "net."
    i = 0
    k = cfg.kernel_size
    x = conv1d(x, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"),
               pad=get_padding(2 * k + 1, mode=cfg.pad_mode))
    if taps is not None:
        taps["stem"] = x
    i += 1
    C = cfg.capacity
    for r, dil in zip(cfg.ratios, cfg.dilations):
        for d in dil:
            if cfg.adain:
                i += 1  # AdaptiveInstanceNormalization is the identity in training
            x = dilated_unit(x, sd, f"{p}{i}.aligned.branches.0.net.", cfg, C, d)
            i += 1
        x = _act(x, sd, f"{p}{i}.", cfg)
        i += 1
        x = conv1d(x, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"), stride=r,
                   pad=get_padding(2 * r, r, mode=cfg.pad_mode))
        i += 1
        C = C * r if cfg.keep_dim else C * 2
        if taps is not None:
            taps[f"down{r}_{C}"] = x
    x = _act(x, sd, f"{p}{i}.", cfg)
    i += 1
    x = conv1d(x, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"),
               pad=get_padding(k, mode=cfg.pad_mode))
    return x


def reparametrize(z: Tensor, eps: Tensor, beta: float = 1.0) -> Tuple[Tensor, Tensor]:
    """VariationalEncoder.reparametrize (rave/blocks.py:725-734) with the noise injected."""
    mean, scale = z.chunk(2, 1)
    std = F.softplus(scale) + 1e-4
    var = std * std
    logvar = torch.log(var)
    zs = eps * std + mean
    kl = (mean * mean + var - logvar - 1).sum(1).mean()
    return zs, beta * kl


def generator_v2(z: Tensor, sd, prefix: str, cfg: ArchConfig,
                 taps: Optional[dict] = None) -> Tensor:
    """GeneratorV2.forward (rave/blocks.py:599-714) without noise module.
    `prefix` e.g. 'decoder.'."""
    p = prefix + "net."
    k = cfg.kernel_size
    ratios = cfg.ratios[::-1]
    dils = cfg.dilations[::-1]
    C = (int(np.prod(ratios)) if cfg.keep_dim else 2 ** len(ratios)) * cfg.capacity
    i = 0
    x = conv1d(z, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"),
               pad=get_padding(k, mode=cfg.pad_mode))
    i += 1
    for r, dil in zip(ratios, dils):
        x = _act(x, sd, f"{p}{i}.", cfg)
        i += 1
        x = conv_transpose1d(x, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"), r, r // 2)
        i += 1
        C = C // r if cfg.keep_dim else C // 2
        if taps is not None:
            taps[f"up{r}_{C}"] = x
        for d in dil:
            if cfg.adain:
                i += 1
            x = dilated_unit(x, sd, f"{p}{i}.aligned.branches.0.net.", cfg, C, d)
            i += 1
    x = _act(x, sd, f"{p}{i}.", cfg)
    i += 1
    x = conv1d(x, wn_weight(sd, f"{p}{i}."), sd.get(f"{p}{i}.bias"),
               pad=get_padding(2 * k + 1, mode=cfg.pad_mode))
    if taps is not None:
        taps["wave"] = x
    if cfg.amplitude_modulation:
        x, amp = x.split(x.shape[1] // 2, 1)
        x = x * torch.sigmoid(amp)
    return torch.tanh(x)


def rave_forward(x: Tensor, sd, cfg: ArchConfig, eps: Tensor,
                 taps: Optional[dict] = None) -> Tensor:
    """RAVE.forward (rave/model.py:267-270) with injected reparametrisation noise."""
    hk = sd["pqmf.hk"]
    x_mb = pqmf_encode(x, hk, cfg.pad_mode)
    z = encoder_v2(x_mb, sd, "encoder.encoder.", cfg)
    zs, _ = reparametrize(z, eps)
    y_mb = generator_v2(zs, sd, "decoder.", cfg)
    if taps is not None:
        taps.update(x_mb=x_mb, z=z, zs=zs, y_mb=y_mb)
    return pqmf_decode(y_mb, hk, cfg.n_channels, cfg.pad_mode)


# ----------------------------------------------------------------------------------
# v2 discriminators (rave/discriminator.py) ----------------------------------------
# ----------------------------------------------------------------------------------


def convnet_1d(x: Tensor, sd, prefix: str, n_layers: int = 4, kernel_size: int = 15,
               stride: int = 4) -> List[Tensor]:
    """ConvNet with conv=nn.Conv1d (rave/discriminator.py:77-119, configs/v1.gin:75-83):
    features are the PRE-activation outputs of every conv, incl. the final 1x1."""
    feats = []
    pad = get_padding(kernel_size, stride, mode="centered")[0]
    for i in range(n_layers):
        q = f"{prefix}net.{2 * i}."
        x = F.conv1d(x, wn_weight(sd, q), sd[q + "bias"], stride, pad)
        feats.append(x)
        x = leaky_relu(x, 0.2)
    q = f"{prefix}net.{2 * n_layers}."
    x = F.conv1d(x, sd[q + "weight"], sd[q + "bias"])
    feats.append(x)
    return feats


def convnet_2d(x: Tensor, sd, prefix: str, n_layers: int = 4, kernel_size: int = 5,
               stride: int = 4) -> List[Tensor]:
    """ConvNet with conv=nn.Conv2d, kernel (5,1) (configs/v2.gin:53-55): stride (s,1),
    padding (get_padding(5, s)[0], 0)."""
    feats = []
    pad = (get_padding(kernel_size, stride, mode="centered")[0], 0)
    for i in range(n_layers):
        q = f"{prefix}net.{2 * i}."
        x = F.conv2d(x, wn_weight(sd, q), sd[q + "bias"], (stride, 1), pad)
        feats.append(x)
        x = leaky_relu(x, 0.2)
    q = f"{prefix}net.{2 * n_layers}."
    x = F.conv2d(x, sd[q + "weight"], sd[q + "bias"])
    feats.append(x)
    return feats


def multi_scale_discriminator(x: Tensor, sd, prefix: str, n: int = 3) -> List[List[Tensor]]:
    """rave/discriminator.py:122-136."""
    out = []
    for i in range(n):
        out.append(convnet_1d(x, sd, f"{prefix}layers.{i}."))
        x = F.avg_pool1d(x, 2)
    return out


def mpd_fold(x: Tensor, n: int) -> Tensor:
    """rave/discriminator.py:192-195."""
    pad = (n - (x.shape[-1] % n)) % n
    x = F.pad(x, (0, pad))
    return x.reshape(*x.shape[:2], -1, n)


def multi_period_discriminator(x: Tensor, sd, prefix: str,
                               periods=(2, 3, 5, 7, 11)) -> List[List[Tensor]]:
    """rave/discriminator.py:174-195."""
    return [convnet_2d(mpd_fold(x, n), sd, f"{prefix}layers.{i}.") for i, n in enumerate(periods)]


def combine_discriminators_v2(x: Tensor, sd, prefix: str = "discriminator.") -> List[List[Tensor]]:
    """CombineDiscriminators[MPD, MSD] (rave/discriminator.py:198-209, configs/v2.gin:65-70)."""
    feats = multi_period_discriminator(x, sd, prefix + "discriminators.0.")
    feats.extend(multi_scale_discriminator(x, sd, prefix + "discriminators.1."))
    return feats


# ----------------------------------------------------------------------------------
# Descript discriminator, MPD part (rave/descript_discriminator.py:30-66) ----------
# ----------------------------------------------------------------------------------


def descript_mpd(x: Tensor, sd, prefix: str, period: int) -> List[Tensor]:
    """Features are POST-activation (each `layer` is Sequential(conv, LeakyReLU(.1)))."""
    t = x.shape[-1]
    x = F.pad(x, (0, period - t % period), mode="reflect")
    x = x.reshape(x.shape[0], x.shape[1], -1, period)
    fmap = []
    strides = [3, 3, 3, 3, 1]
    for i, s in enumerate(strides):
        q = f"{prefix}convs.{i}.0."
        x = F.conv2d(x, wn_weight(sd, q), sd[q + "bias"], (s, 1), (2, 0))
        x = leaky_relu(x, 0.1)
        fmap.append(x)
    q = f"{prefix}conv_post."
    x = F.conv2d(x, wn_weight(sd, q), sd[q + "bias"], 1, (1, 0))
    fmap.append(x)
    return fmap


DESCRIPT_BANDS = [(0.0, 0.1), (0.1, 0.25), (0.25, 0.5), (0.5, 0.75), (0.75, 1.0)]


def descript_mrd(x: Tensor, sd, prefix: str, window_length: int, bands=DESCRIPT_BANDS) -> List[Tensor]:
    """MRD (rave/descript_discriminator.py:118-184): complex STFT (torchaudio Spectrogram, hann, hop = wl/4, centred,
    power=None) as [b, (c re/im), t, f], split into 5 frequency bands, each through its own stack of (3,9)/(3,3)
    weight-normed Conv2d + LeakyReLU(.1) (stride 2 along frequency in layers 1-3); the band outputs are concatenated
    along frequency for conv_post.  Features are POST-activation, 5 x 5 + 1 = 26 of them."""
    n_fft = window_length // 2 + 1
    bnd = [(int(b[0] * n_fft), int(b[1] * n_fft)) for b in bands]
    B, C, T = x.shape
    win = torch.hann_window(window_length, dtype=x.dtype, device=x.device)
    s = torch.stft(x.reshape(B * C, T), window_length, hop_length=int(0.25 * window_length),
                   win_length=window_length, window=win, center=True, pad_mode="reflect", normalized=False,
                   onesided=True, return_complex=True)                       # [B*C, f, t]
    s = torch.view_as_real(s).reshape(B, C, s.shape[-2], s.shape[-1], 2)     # b c f t p
    xs = s.permute(0, 1, 4, 3, 2).reshape(B, 2 * C, s.shape[3], s.shape[2])  # b (c p) t f
    fmap, outs = [], []
    strides = [(1, 1), (1, 2), (1, 2), (1, 2), (1, 1)]
    pads = [(1, 4)] * 4 + [(1, 1)]
    for bi, (lo, hi) in enumerate(bnd):
        band = xs[..., lo:hi]
        for li in range(5):
            q = f"{prefix}band_convs.{bi}.{li}.0."
            band = leaky_relu(F.conv2d(band, wn_weight(sd, q), sd[q + "bias"], strides[li], pads[li]), 0.1)
            fmap.append(band)
        outs.append(band)
    q = f"{prefix}conv_post."
    fmap.append(F.conv2d(torch.cat(outs, -1), wn_weight(sd, q), sd[q + "bias"], 1, (1, 1)))
    return fmap


def descript_discriminator(x: Tensor, sd, prefix: str = "discriminator.", periods=(2, 3, 5, 7, 11),
                           fft_sizes=(2048, 1024, 512)) -> List[List[Tensor]]:
    """DescriptDiscriminator.forward (rave/descript_discriminator.py:187-217): preprocess, 5 MPDs, 3 MRDs."""
    x = descript_preprocess(x)
    out = [descript_mpd(x, sd, f"{prefix}discriminators.{i}.", p) for i, p in enumerate(periods)]
    n = len(periods)
    out += [descript_mrd(x, sd, f"{prefix}discriminators.{n + i}.", w) for i, w in enumerate(fft_sizes)]
    return out


def descript_preprocess(y: Tensor) -> Tensor:
    """DescriptDiscriminator.preprocess (rave/descript_discriminator.py:207-212)."""
    y = y - y.mean(dim=-1, keepdims=True)
    return 0.8 * y / (y.abs().max(dim=-1, keepdim=True)[0] + 1e-9)


# ----------------------------------------------------------------------------------
# Losses (rave/core.py) -- "next" rows ---------------------
# ----------------------------------------------------------------------------------


def stft_mag(x: Tensor, n_fft: int) -> Tensor:
    """torchaudio.transforms.Spectrogram(n_fft, win_length=n_fft, hop=n_fft//4, power=None)
    followed by abs (rave/core.py:269-319): hann (periodic) window, centred, reflect pad."""
    win = torch.hann_window(n_fft, dtype=x.dtype, device=x.device)
    s = torch.stft(x, n_fft, hop_length=n_fft // 4, win_length=n_fft, window=win,
                   center=True, pad_mode="reflect", normalized=False, onesided=True,
                   return_complex=True)
    return s.abs()


def mean_difference(target: Tensor, value: Tensor, norm: str = "L1", relative: bool = False):
    """rave/core.py:236-252."""
    diff = target - value
    if norm == "L1":
        diff = diff.abs().mean()
        if relative:
            diff = diff / target.abs().mean()
        return diff
    diff = (diff * diff).mean()
    if relative:
        diff = diff / (target * target).mean()
    return diff


def audio_distance_v1(x: Tensor, y: Tensor, scales=(2048, 1024, 512, 256, 128),
                      log_epsilon: float = 1e-7) -> Tensor:
    """AudioDistanceV1.forward (rave/core.py:322-344)."""
    x = x.reshape(-1, x.shape[-1])
    y = y.reshape(-1, y.shape[-1])
    distance = 0.0
    for s in scales:
        sx, sy = stft_mag(x, s), stft_mag(y, s)
        lin = mean_difference(sx, sy, "L2", relative=True)
        log = mean_difference(torch.log(sx + log_epsilon), torch.log(sy + log_epsilon), "L1")
        distance = distance + lin + log
    return distance


def hinge_gan(score_real: Tensor, score_fake: Tensor):
    """rave/core.py:151-155."""
    loss_dis = (torch.relu(1 - score_real) + torch.relu(1 + score_fake)).mean()
    return loss_dis, -score_fake.mean()


def gan_losses(features: List[List[Tensor]], num_skipped_features: int = 1,
               relative: bool = True):
    """The discrimination block of RAVE.training_step (rave/model.py:348-379) given the
    discriminator output on cat([x, y])."""
    fm = 0.0
    loss_dis = 0.0
    loss_adv = 0.0
    for scale in features:
        real = [f[: f.shape[0] // 2] for f in scale]
        fake = [f[f.shape[0] // 2:] for f in scale]
        r, k = real[num_skipped_features:], fake[num_skipped_features:]
        fm = fm + sum(mean_difference(a, b, "L1", relative) for a, b in zip(r, k)) / len(r)
        d, a = hinge_gan(real[-1], fake[-1])
        loss_dis = loss_dis + d
        loss_adv = loss_adv + a
    return fm / len(features), loss_dis, loss_adv


def v2_config(**kw) -> ArchConfig:
    """configs/v2.gin:12-50."""
    return ArchConfig(**kw)


def v2_small_config(**kw) -> ArchConfig:
    """configs/v2_small.gin:12-21 (noise module handled by the caller)."""
    base = dict(capacity=48, ratios=(4, 2, 2, 2))
    base.update(kw)
    return ArchConfig(**base)


# ----------------------------------------------------------------------------------
# CPU training-step arithmetic (bench.py cpu_baseline / --impl reference) ----------
# ----------------------------------------------------------------------------------


def valid_signal_crop(x: Tensor, left_rf: int, right_rf: int) -> Tensor:
    """rave/core.py:220-225 (the `-right_rf // dim` floor division of a negative number included)."""
    dim = x.shape[1]
    x = x[..., left_rf // dim:]
    if right_rf:
        x = x[..., :-right_rf // dim]
    return x


def train_step_losses(x: Tensor, sd, cfg: ArchConfig, eps: Tensor, warmed_up: bool = True,
                      beta: float = 1.0, fm_weight: float = 20.0, receptive_field=(0, 0),
                      return_parts: bool = False):
    """Forward arithmetic of RAVE.training_step (rave/model.py:292-399) for the v2 family,
    returning (loss_gen_total, loss_dis[, the logged loss_gen terms]).  Quirk D1 (weights applied
    twice, rave/model.py:397,410-411) included; `receptive_field` = the buffer valid_signal_crop reads
    (rave/model.py:322-330).  Pinned against the reference's own training_step by
    oracle/make_golden.py::golden_training_step (tests/golden/training_step_v2_tiny.pt)."""
    hk = sd["pqmf.hk"]
    x_mb = pqmf_encode(x, hk, cfg.pad_mode)
    z = encoder_v2(x_mb, sd, "encoder.encoder.", cfg)
    if warmed_up:
        z = z.detach()                                   # blocks.py:743-744
    zs, reg = reparametrize(z, eps)
    y_mb = generator_v2(zs, sd, "decoder.", cfg)
    y = pqmf_decode(y_mb, hk, cfg.n_channels, cfg.pad_mode)
    y = y[..., :x.shape[-1]]
    y_mb = y_mb[..., :x_mb.shape[-1]]
    x_mb_c, y_mb_c = x_mb, y_mb
    if receptive_field[0] + receptive_field[1]:
        x_mb_c = valid_signal_crop(x_mb, *receptive_field)
        y_mb_c = valid_signal_crop(y_mb, *receptive_field)
    losses = {
        "multiband_spectral_distance": audio_distance_v1(x_mb_c, y_mb_c),
        "fullband_spectral_distance": audio_distance_v1(x, y),
        "regularization": reg * beta,
    }
    loss_dis = torch.zeros(())
    if warmed_up:
        feats = combine_discriminators_v2(torch.cat([x, y], 0), sd)
        fm, loss_dis, loss_adv = gan_losses(feats, 1, True)
        losses["feature_matching"] = fm_weight * fm
        losses["adversarial"] = loss_adv
    weights = {"feature_matching": fm_weight}
    total = sum(v * weights.get(k, 1.0) for k, v in losses.items())
    if return_parts:
        return total, loss_dis, losses
    return total, loss_dis


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, betas=(0.5, 0.9),
              eps: float = 1e-8):
    """torch.optim.Adam (no weight decay / amsgrad) as rave/model.py:226-236 configures it; returns the
    new (p, m, v).  `step` is the 1-based count after this update."""
    b1, b2 = betas
    m = m * b1 + g * (1 - b1)
    v = v * b2 + g * g * (1 - b2)
    denom = (v.sqrt() / math.sqrt(1 - b2 ** step)) + eps
    return p - (lr / (1 - b1 ** step)) * m / denom, m, v


def train_step_cpu(x: Tensor, sd, cfg: ArchConfig, eps: Tensor, dis_step: bool, warmed_up: bool = True,
                   receptive_field=(0, 0), return_named: bool = False):
    """One forward+backward of the reference's training step on CPU via autograd (no optimiser
    state: the timing baseline counts the same fwd+bwd work the GPU arm does).  Gradients go to the group
    the reference steps: discriminator (D-step), encoder+decoder (G-step; the encoder's are None in phase 2)."""
    params = {k: v.clone().requires_grad_(v.is_floating_point() and not k.startswith("pqmf."))
              for k, v in sd.items()}
    total, loss_dis = train_step_losses(x, params, cfg, eps, warmed_up, receptive_field=receptive_field)
    if dis_step:
        names = [k for k, v in params.items() if k.startswith("discriminator.") and v.requires_grad]
        grads = torch.autograd.grad(loss_dis, [params[k] for k in names], allow_unused=True)
    else:
        names = [k for k, v in params.items() if (k.startswith("decoder.") or (k.startswith("encoder.")
                                                                               and not warmed_up))
                 and v.requires_grad]
        grads = torch.autograd.grad(total, [params[k] for k in names], allow_unused=True)
    if return_named:
        return total.detach(), loss_dis.detach(), dict(zip(names, grads))
    return total.detach(), loss_dis.detach(), grads


# ----------------------------------------------------------------------------------
# NoiseGeneratorV2 (rave/blocks.py:243-292; helpers rave/core.py:20-21, 48-81) ------
# ----------------------------------------------------------------------------------


def mod_sigmoid(x: Tensor) -> Tensor:
    """rave/core.py:20-21."""
    return 2 * torch.sigmoid(x) ** 2.3 + 1e-7


def amp_to_impulse_response(amp: Tensor, target_size: int) -> Tensor:
    """rave/core.py:48-69: zero-phase amplitudes -> irfft -> centre -> hann -> pad / crop -> un-centre."""
    amp = torch.view_as_complex(torch.stack([amp, torch.zeros_like(amp)], -1))
    amp = torch.fft.irfft(amp)
    filter_size = amp.shape[-1]
    amp = torch.roll(amp, filter_size // 2, -1)
    amp = amp * torch.hann_window(filter_size, dtype=amp.dtype, device=amp.device)
    amp = F.pad(amp, (0, int(target_size) - int(filter_size)))
    return torch.roll(amp, -filter_size // 2, -1)


def fft_convolve(signal: Tensor, kernel: Tensor) -> Tensor:
    """rave/core.py:71-81."""
    signal = F.pad(signal, (0, signal.shape[-1]))
    kernel = F.pad(kernel, (kernel.shape[-1], 0))
    output = torch.fft.irfft(torch.fft.rfft(signal) * torch.fft.rfft(kernel))
    return output[..., output.shape[-1] // 2:]


def noise_generator_v2(x: Tensor, sd, prefix: str, ratios=(2, 2, 2), data_size: int = 16, n_channels: int = 1,
                       noise: Optional[Tensor] = None, slope: float = 0.2) -> Tensor:
    """rave/blocks.py:243-292 with the uniform noise injected (`noise` [B, T', C, target] in [-1, 1); the reference
    draws it with torch.rand_like).  Convs: cc.Conv1d(k = 2r, stride r, padding (r, 0)), LeakyReLU(.2) between them
    (configs/v2_small.gin:42-57), plain weights + bias (`normalization` is not applied to this branch)."""
    h = x
    for i, r in enumerate(ratios):
        idx = 2 * i                       # nn.Sequential: conv, act, conv, act, conv
        if i:
            h = leaky_relu(h, slope)
        h = conv1d(h, sd[f"{prefix}net.{idx}.weight"], sd.get(f"{prefix}net.{idx}.bias"), r, 1, (r, 0))
    amp = mod_sigmoid(h - 5)
    amp = amp.permute(0, 2, 1)
    amp = amp.reshape(amp.shape[0], amp.shape[1], n_channels * data_size, -1)
    target = 1
    for r in ratios:
        target *= r
    ir = amp_to_impulse_response(amp, target)
    if noise is None:
        noise = torch.rand_like(ir) * 2 - 1
    out = fft_convolve(noise, ir).permute(0, 2, 1, 3)
    return out.reshape(out.shape[0], out.shape[1], -1)


# ----------------------------------------------------------------------------------
# v1 architecture (rave/blocks.py:48-240, 322-503; configs/v1.gin) -------------------
# ----------------------------------------------------------------------------------

V1_DILATIONS = ((1, 1), (3, 1), (5, 1))        # configs/v1.gin:63-65 (ResidualStack.dilations_list, kernel_sizes [3])


def encoder_v1(x: Tensor, sd, prefix: str, ratios=(4, 4, 4, 2), n_out: int = 2, mode: str = "centered",
               training: bool = True) -> Tensor:
    """Encoder.forward (rave/blocks.py:426-503) with sample_norm=False, repeat_layers=1 (v1.gin:44-50): conv k7, then per
    ratio BatchNorm1d -> LeakyReLU(.2) -> Conv1d(k = 2r+1, stride r), LeakyReLU, grouped Conv1d(k5, groups = n_out).
    BatchNorm uses batch statistics in training mode (nn.BatchNorm1d defaults: eps 1e-5)."""
    p = prefix + "net."
    h = conv1d(x, sd[p + "0.weight"], sd.get(p + "0.bias"), pad=get_padding(7, mode=mode))
    i = 1
    for r in ratios:
        h = F.batch_norm(h, sd[f"{p}{i}.running_mean"].clone(), sd[f"{p}{i}.running_var"].clone(),
                         sd[f"{p}{i}.weight"], sd[f"{p}{i}.bias"], training, 0.1, 1e-5)
        h = leaky_relu(h, 0.2)
        h = conv1d(h, sd[f"{p}{i + 2}.weight"], sd.get(f"{p}{i + 2}.bias"), r, 1, get_padding(2 * r + 1, r, mode=mode))
        i += 3
    h = leaky_relu(h, 0.2)
    w, b = sd[f"{p}{i + 1}.weight"], sd.get(f"{p}{i + 1}.bias")
    return F.conv1d(F.pad(h, get_padding(5, mode=mode)), w, b, 1, 0, 1, n_out)


def residual_stack_v1(x: Tensor, sd, prefix: str, mode: str = "centered") -> Tensor:
    """ResidualStack (rave/blocks.py:144-160) with kernel_sizes [3]: one ResidualBlock of three ResidualLayers
    x + conv(act(conv(act(x)))) (48-80, 115-141); the single branch is stacked and summed (158-159)."""
    h = x
    for li, dil in enumerate(V1_DILATIONS):
        q = f"{prefix}net.branches.0.net.{li}.net.aligned.branches.0."
        y = h
        for j, d in enumerate(dil):
            y = leaky_relu(y, 0.2)
            y = conv1d(y, wn_weight(sd, f"{q}{2 * j + 1}."), sd.get(f"{q}{2 * j + 1}.bias"), 1, d,
                       get_padding(3, dilation=d, mode=mode))
        h = h + y
    return h


def noise_generator_v1(x: Tensor, sd, prefix: str, ratios=(4, 4, 4), data_size: int = 16,
                       noise: Optional[Tensor] = None, mode: str = "centered") -> Tensor:
    """NoiseGenerator.forward (rave/blocks.py:195-240): strided convs -> mod_sigmoid(. - 5) -> impulse responses ->
    FFT convolution with uniform noise (injected, like noise_generator_v2)."""
    h = x
    for i, r in enumerate(ratios):
        h = conv1d(h, sd[f"{prefix}net.{2 * i}.weight"], sd.get(f"{prefix}net.{2 * i}.bias"), r, 1,
                   get_padding(3, r, mode=mode))
        if i != len(ratios) - 1:
            h = leaky_relu(h, 0.2)
    amp = mod_sigmoid(h - 5).permute(0, 2, 1)
    amp = amp.reshape(amp.shape[0], amp.shape[1], data_size, -1)
    target = 1
    for r in ratios:
        target *= r
    ir = amp_to_impulse_response(amp, target)
    if noise is None:
        noise = torch.rand_like(ir) * 2 - 1
    out = fft_convolve(noise, ir).permute(0, 2, 1, 3)
    return out.reshape(out.shape[0], out.shape[1], -1)


def generator_v1(z: Tensor, sd, prefix: str, ratios=(4, 4, 4, 2), data_size: int = 16, loud_stride: int = 1,
                 use_noise: bool = True, warmed_up: bool = False, noise: Optional[Tensor] = None,
                 mode: str = "centered") -> Tensor:
    """Generator.forward (rave/blocks.py:322-423): conv k7, per ratio UpsampleLayer (act -> ConvTranspose1d(2r, r, r//2),
    163-192) + ResidualStack, then tanh(waveform) * mod_sigmoid(loudness) (+ noise once warmed up)."""
    p = prefix + "net."
    h = conv1d(z, wn_weight(sd, p + "0."), sd.get(p + "0.bias"), pad=get_padding(7, mode=mode))
    i = 1
    for r in ratios:
        h = leaky_relu(h, 0.2)
        if r > 1:
            h = conv_transpose1d(h, wn_weight(sd, f"{p}{i}.net.1."), sd.get(f"{p}{i}.net.1.bias"), r, r // 2)
        else:
            h = conv1d(h, wn_weight(sd, f"{p}{i}.net.1."), sd.get(f"{p}{i}.net.1.bias"), pad=get_padding(3, mode=mode))
        h = residual_stack_v1(h, sd, f"{p}{i + 1}.", mode)
        i += 2
    q = prefix + "synth.branches."
    wave = conv1d(h, wn_weight(sd, q + "0."), sd.get(q + "0.bias"), pad=get_padding(7, mode=mode))
    loud = conv1d(h, wn_weight(sd, q + "1."), sd.get(q + "1.bias"), loud_stride, 1,
                  get_padding(2 * loud_stride + 1, loud_stride, mode=mode))
    if loud_stride != 1:
        loud = loud.repeat_interleave(loud_stride)
    loud = loud.reshape(h.shape[0], 1, -1)
    wave = torch.tanh(wave) * mod_sigmoid(loud)
    if warmed_up and use_noise:
        wave = wave + noise_generator_v1(h, sd, q + "2.", data_size=data_size, noise=noise, mode=mode)
    return wave
