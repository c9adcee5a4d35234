"""CPU: the oracle restatement reproduces every committed golden tensor (the goldens were
produced by executing the unmodified reference, oracle/make_golden.py)."""
import os

import pytest
import torch

from oracle import rave_oracle as O
from tests.conftest import GOLDEN, rel_l2


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def test_pqmf_design_and_operators():
    g = load("pqmf.pt")
    h, hk = O.pqmf_design(100, 16)
    assert torch.equal(h, g["h"]) and torch.equal(hk, g["hk"])
    assert h.shape == (377,) and hk.shape == (16, 512)
    hkf, hki = O.cached_pqmf_weights(hk)
    assert torch.equal(hkf, g["forward_conv.weight"]) and torch.equal(hki, g["inverse_conv.weight"])
    for mode in ("centered", "causal"):
        x, y, xr = g[mode]["x"], g[mode]["y"], g[mode]["xr"]
        assert rel_l2(O.pqmf_analysis(x, hk, mode), y) < 1e-7
        assert rel_l2(O.pqmf_synthesis(y, hk, mode), xr) < 1e-7


def test_pqmf_three_formulations_agree():
    """cached / polyphase / classic analysis agree; CachedPQMF.inverse == PQMF.inverse
    delayed by 16 samples."""
    g = load("pqmf.pt")
    hk = g["hk"]
    x, y = g["centered"]["x"], g["centered"]["y"]
    assert rel_l2(O.polyphase_forward(x, hk), y) < 4e-7
    assert rel_l2(O.classic_forward(x, hk), y) < 4e-7
    inv_poly = O.polyphase_inverse(y, hk)
    assert rel_l2(inv_poly, g["polyphase_inverse"]) < 1e-7
    inv_cached = O.pqmf_synthesis(y, hk)
    assert (inv_cached[..., 16:] - inv_poly[..., :-16]).abs().max() < 1e-5


def test_pqmf_roundtrip_config1():
    """BASELINE config 1: 16-band analysis->synthesis on 1 x 131072 @48 kHz, CPU.  Near-perfect
    reconstruction only: rel-L2 ~ 1.0e-3 (100 dB pseudo-QMF), 16-sample delay."""
    g = load("pqmf.pt")
    hk = g["hk"]
    gen = torch.Generator().manual_seed(7)
    x = (0.5 * torch.randn(1, 1, 131072, generator=gen)).clamp(-1, 1)
    xr = O.pqmf_synthesis(O.pqmf_analysis(x, hk), hk)
    assert xr.shape == x.shape
    r = rel_l2(xr[..., 16 + 1024:-1024], x[..., 1024:-1024 - 16])
    assert abs(r - g["roundtrip_rel_l2"]) < 1e-6
    assert 0.9e-3 < r < 1.1e-3


@pytest.mark.parametrize("name", ["v2_tiny", "v2_tiny_causal", "v3_tiny", "v2_small_tiny"])
def test_autoencoder_golden(name):
    g = load(f"autoencoder_{name}.pt")
    cfg = O.ArchConfig(**{k: v for k, v in g["cfg"].items()})
    taps = {}
    y = O.rave_forward(g["x"], g["state_dict"], cfg, g["eps"], taps)
    assert rel_l2(taps["x_mb"], g["x_mb"]) < 1e-7
    assert rel_l2(taps["z"], g["z"]) < 2e-6
    assert rel_l2(taps["y_mb"], g["y_mb"]) < 2e-6
    assert rel_l2(y, g["y"]) < 2e-6
    assert y.shape == g["x"].shape                      # tests/test_configs.py:69 contract


def test_discriminator_golden():
    g = load("discriminator_v2.pt")
    feats = O.combine_discriminators_v2(g["x"], g["state_dict"])
    assert len(feats) == 8 and all(len(f) == 5 for f in feats)
    for fa, fb in zip(feats, g["features"]):
        for a, b in zip(fa, fb):
            assert a.shape == b.shape and rel_l2(a, b) < 1e-6
    fm, ld, la = O.gan_losses(feats, 1, True)
    assert rel_l2(fm, g["fm"]) < 1e-6 and rel_l2(ld, g["loss_dis"]) < 1e-6 and rel_l2(la, g["loss_adv"]) < 1e-5


def test_audio_distance_golden():
    g = load("audio_distance.pt")
    assert rel_l2(O.audio_distance_v1(g["x"], g["y"]), g["distance"]) < 1e-6


@pytest.mark.reference
def test_oracle_matches_live_reference():
    """Build container only: re-run the unmodified reference and compare (a fresh pin)."""
    from oracle.ref_loader import load_reference
    R = load_reference()
    p = R.pqmf.CachedPQMF(attenuation=100, n_band=16)
    x = torch.randn(1, 1, 4096)
    assert torch.equal(O.pqmf_analysis(x, p.hk), p(x))
    y = p(x)
    assert torch.equal(O.pqmf_synthesis(y, p.hk), p.inverse(y))


@pytest.mark.reference
def test_descript_mpd_oracle_matches_live_reference():
    """Descript MPD (rave/descript_discriminator.py:30-66): too large for a committed fixture (1024-ch
    layers), so the oracle restatement is pinned against the live reference in the build container."""
    from oracle.ref_loader import load_reference
    R = load_reference()
    torch.manual_seed(0)
    mpd = R.descript_discriminator.MPD(3)
    x = torch.randn(2, 1, 999)
    want = mpd(x)
    sd = {k: v.detach() for k, v in mpd.state_dict().items()}
    got = O.descript_mpd(x, sd, "", 3)
    assert len(got) == len(want) == 6
    for a, b in zip(got, want):
        assert a.shape == b.shape and rel_l2(a, b) < 1e-6
    y = torch.randn(2, 1, 500)
    dd = R.descript_discriminator.DescriptDiscriminator()
    assert torch.equal(O.descript_preprocess(y), dd.preprocess(y))
