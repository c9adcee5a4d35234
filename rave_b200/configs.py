"""gin-free instantiation of the reference's configurations (gin-config is not installable
here; SURVEY.md App. B.1 lists the effective bindings with file:line provenance).

`build_rave("v2", sampling_rate=48000)` returns the same module tree (same `state_dict` keys and
shapes) that `scripts/train.py --config v2 --override SAMPLING_RATE=48000` would build through
`rave.RAVE()` (scripts/train.py:139-159).
"""
from contextlib import contextmanager
from functools import partial

import torch.nn as nn

from . import blocks, cc, core, discriminator, pqmf, quantization
from .model import RAVE

V2_DILATIONS = [[1, 3, 9], [1, 3, 9], [1, 3, 9], [1, 3]]          # configs/v2.gin:13-18

ARCH = {
    # name: capacity, ratios, activation, adain, discriminator, update_discriminator_every, phase_1
    "v2": dict(capacity=96, ratios=[4, 4, 4, 2], activation="leaky", adain=False,     # v2.gin:12-21
               disc="v2", update_discriminator_every=4, phase_1_duration=1000000),
    "v2_small": dict(capacity=48, ratios=[4, 2, 2, 2], activation="leaky", adain=False,  # v2_small.gin:12-21
                     disc="v2", update_discriminator_every=2, phase_1_duration=1000000),
    "v3": dict(capacity=96, ratios=[4, 4, 4, 2], activation="snake", adain=True,       # v3.gin:3-13
               disc="descript", update_discriminator_every=4, phase_1_duration=1000000),
    # discrete.gin:13-49 (EnCodec-style RVQ head; generator latent = 128 + 128 noise channels)
    "discrete": dict(capacity=96, ratios=[4, 4, 2, 2], activation="leaky", adain=False, disc="v2",
                     update_discriminator_every=4, phase_1_duration=200000, discrete=True,
                     noise_augmentation=128, log_epsilon=1.0, num_skipped_features=0),
}


def _activation_factory(kind):
    if kind == "snake":
        return lambda dim: blocks.Snake(dim)          # configs/snake.gin:5-23
    return lambda dim: nn.LeakyReLU(.2)


def make_autoencoder(name="v2", capacity=None, latent_size=128, n_band=16, n_channels=1,
                     padding_mode="centered", ratios=None, activation=None, adain=None, with_noise=False):
    """(pqmf, encoder, decoder) factories -> constructed modules for one architecture."""
    a = ARCH[name]
    capacity = capacity or a["capacity"]
    ratios = ratios or a["ratios"]
    act = _activation_factory(activation or a["activation"])
    use_adain = a["adain"] if adain is None else adain
    adain_f = (lambda dim: blocks.AdaptiveInstanceNormalization(dim)) if use_adain else None
    with cc.configure(conv_bias=False, padding_mode=padding_mode):       # v1.gin:33-34, causal.gin:5
        pq = pqmf.CachedPQMF(attenuation=100, n_band=n_band, n_channels=n_channels)  # v1.gin:37-39
        enc = blocks.VariationalEncoder(                                  # v2.gin:30-40
            partial(blocks.EncoderV2, data_size=n_band, capacity=capacity, ratios=ratios,
                    latent_size=latent_size, n_out=2, kernel_size=3, dilations=V2_DILATIONS,
                    activation=act, adain=adain_f),
            n_channels=n_channels)
        noise = None
        if name == "v2_small" and with_noise:                                  # v2_small.gin:42-57
            noise = partial(blocks.NoiseGeneratorV2, hidden_size=64, data_size=n_band, ratios=[2, 2, 2],
                            noise_bands=32, activation=act)
        dec = blocks.GeneratorV2(data_size=n_band, capacity=capacity, ratios=ratios,  # v2.gin:43-50
                                 latent_size=latent_size, kernel_size=3, dilations=V2_DILATIONS,
                                 amplitude_modulation=True, activation=act, adain=adain_f,
                                 n_channels=n_channels, noise_module=noise)
    return pq, enc, dec


def make_discriminator_v2(capacity=96, n_channels=1):
    """CombineDiscriminators[MPD(2,3,5,7,11), MSD(3)] (configs/v2.gin:53-75, v1.gin:75-88)."""
    periods_net = partial(discriminator.ConvNet, out_size=1, capacity=capacity, n_layers=4, stride=4,
                          conv=nn.Conv2d, kernel_size=(5, 1))
    scales_net = partial(discriminator.ConvNet, out_size=1, capacity=capacity, n_layers=4, stride=4,
                         conv=nn.Conv1d, kernel_size=15)
    return discriminator.CombineDiscriminators([
        partial(discriminator.MultiPeriodDiscriminator, periods=[2, 3, 5, 7, 11], convnet=periods_net),
        partial(discriminator.MultiScaleDiscriminator, n_discriminators=3, convnet=scales_net),
    ], n_channels=n_channels)


def build_rave(name="v2", sampling_rate=48000, capacity=None, latent_size=128, n_channels=1,
               padding_mode="centered", phase_1_duration=None, disc_capacity=None, ratios=None):
    """The full `RAVE` model of a named configuration."""
    a = ARCH[name]
    cap = capacity or a["capacity"]
    act = _activation_factory(a["activation"])
    adain_f = (lambda dim: blocks.AdaptiveInstanceNormalization(dim)) if a["adain"] else None
    rat = ratios or a["ratios"]
    stft = partial(core.MultiScaleSTFT, scales=[2048, 1024, 512, 256, 128],        # v1.gin:21-28
                   sample_rate=sampling_rate, magnitude=True)
    distance = partial(core.AudioDistanceV1, multiscale_stft=stft, log_epsilon=a.get("log_epsilon", 1e-7))
    if a["disc"] == "v2":
        disc = lambda n_channels=1: make_discriminator_v2(disc_capacity or cap, n_channels)
    else:
        from .descript_discriminator import DescriptDiscriminator
        disc = lambda n_channels=1: DescriptDiscriminator(n_channels=n_channels)
    noise_aug = a.get("noise_augmentation", 0)
    if a.get("discrete"):
        encoder = partial(blocks.DiscreteEncoder,                                # discrete.gin:26-38
                          encoder_cls=partial(blocks.EncoderV2, data_size=16, capacity=cap, ratios=rat,
                                              latent_size=latent_size, n_out=1, kernel_size=3,
                                              dilations=V2_DILATIONS, activation=act, adain=adain_f),
                          vq_cls=partial(quantization.ResidualVectorQuantization, num_quantizers=16,
                                         dim=latent_size, codebook_size=1024),
                          num_quantizers=16, noise_augmentation=noise_aug)
    else:
        encoder = partial(blocks.VariationalEncoder,
                          partial(blocks.EncoderV2, data_size=16, capacity=cap, ratios=rat,
                                  latent_size=latent_size, n_out=2, kernel_size=3,
                                  dilations=V2_DILATIONS, activation=act, adain=adain_f))
    noise = None
    if name == "v2_small":                                                       # v2_small.gin:42-57
        noise = partial(blocks.NoiseGeneratorV2, hidden_size=64, data_size=16, ratios=[2, 2, 2],
                        noise_bands=32, activation=act)
    with cc.configure(conv_bias=False, padding_mode=padding_mode):
        model = RAVE(
            latent_size=latent_size, sampling_rate=sampling_rate,
            pqmf=partial(pqmf.CachedPQMF, attenuation=100, n_band=16),
            encoder=encoder,
            decoder=partial(blocks.GeneratorV2, data_size=16, capacity=cap, ratios=rat,
                            latent_size=core.get_augmented_latent_size(latent_size, noise_aug), kernel_size=3,
                            dilations=V2_DILATIONS, amplitude_modulation=True, activation=act, adain=adain_f,
                            noise_module=noise),
            discriminator=disc,
            phase_1_duration=phase_1_duration if phase_1_duration is not None else a["phase_1_duration"],
            gan_loss=core.hinge_gan, valid_signal_crop=True,                  # v2.gin:81-83
            feature_matching_fun=partial(core.mean_difference, norm="L1", relative=True),
            num_skipped_features=a.get("num_skipped_features", 1),
            audio_distance=distance, multiband_audio_distance=distance,
            weights={"feature_matching": 20},                                 # v2.gin:87-89
            update_discriminator_every=a["update_discriminator_every"], n_channels=n_channels)
    return model
