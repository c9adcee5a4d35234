"""CPU: host logic of the product (no device compute): C-ABI symbols, module surface /
state_dict contract, padding rules, PQMF filter design, loud failure without a GPU."""
import ctypes
import os
import re

import pytest
import torch
import torch.nn as nn

import rave_b200
from rave_b200 import _lib, blocks, cc, configs, pqmf
from tests.conftest import GOLDEN, ROOT


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rave_b200.h")).read()
    declared = set(re.findall(r"\b(rave_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/rave_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().rave_b200_version() >= 100


def test_no_cpu_fallback():
    conv = cc.Conv1d(4, 4, 3, padding=cc.get_padding(3))
    with pytest.raises(_lib.RaveB200Error):
        conv(torch.randn(1, 4, 16))
    with cc.configure():
        p = pqmf.CachedPQMF(100, 16)
    with pytest.raises(_lib.RaveB200Error):
        p(torch.randn(1, 1, 1024))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "rave_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_get_padding_rules():
    # SURVEY App. A
    assert cc.get_padding(1) == (0, 0)
    assert cc.get_padding(7) == (3, 3)
    assert cc.get_padding(3, dilation=9) == (9, 9)
    assert cc.get_padding(8, 4) == (3, 4)
    assert cc.get_padding(4, 2) == (1, 2)
    assert cc.get_padding(513) == (256, 256)
    assert cc.get_padding(33) == (16, 16)
    assert cc.get_padding(513, mode="causal") == (512, 0)
    assert cc.get_padding(33, mode="causal") == (32, 0)
    assert cc.get_padding(8, 4, mode="causal") == (7, 0)
    with cc.configure(padding_mode="causal"):
        assert cc.get_padding(3) == (2, 0)
    assert cc.get_padding(3) == (1, 1)


def test_pqmf_design_matches_reference_golden():
    g = torch.load(os.path.join(GOLDEN, "pqmf.pt"), weights_only=False)
    p = pqmf.CachedPQMF(attenuation=100, n_band=16)
    assert torch.equal(p.h, g["h"]) and torch.equal(p.hk, g["hk"])
    assert torch.equal(p.forward_conv.weight.data, g["forward_conv.weight"])
    assert torch.equal(p.inverse_conv.weight.data, g["inverse_conv.weight"])
    assert p.forward_conv._pad == (256, 256) and p.inverse_conv._pad == (16, 16)
    assert set(p.state_dict()) == {"hk", "h", "forward_conv.weight", "inverse_conv.weight"}
    with cc.configure(padding_mode="causal"):
        pc = pqmf.CachedPQMF(attenuation=100, n_band=16)
    assert pc.forward_conv._pad == (512, 0) and pc.inverse_conv._pad == (32, 0)


def test_pqmf_adjoint_tables_are_exact_adjoints():
    """The backward kernels are the forward kernels with re-indexed taps (ops.py); check the
    re-indexing against autograd of a dense CPU restatement."""
    from oracle import rave_oracle as O
    p = pqmf.CachedPQMF(attenuation=100, n_band=16)
    t = pqmf._build_tables(p.forward_conv.weight[:, 0, :].detach(), 256, 256,
                           p.inverse_conv.weight.detach(), 16)
    T = 1024
    x = torch.randn(1, 1, T, dtype=torch.float64, requires_grad=True)
    hk = p.hk.double()
    y = O.pqmf_analysis(x, hk)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    # adjoint as a synthesis:  out[16t+15-m] = sum_c sum_j w'[m][c][j] s(c,tau) gy[c][tau], tau=t+j-P
    wb = t["dense"]["taps_bwd"].double()
    P = t["taps_bwd_pad"]
    s = O.reverse_half(gy)
    conv = torch.nn.functional.conv1d(torch.nn.functional.pad(s, (P, 32 - P)), wb)
    out = conv.flip(1).permute(0, 2, 1).reshape(1, 1, -1)
    assert torch.allclose(out, gx, atol=1e-12)
    # adjoint of synthesis as an analysis
    yb = torch.randn(1, 16, 64, dtype=torch.float64, requires_grad=True)
    xo = O.pqmf_synthesis(yb, hk)
    go = torch.randn_like(xo)
    (gyb,) = torch.autograd.grad(xo, yb, go)
    tb = t["dense"]["w_bwd"].double()
    pad = t["w_bwd_pad"]
    an = torch.nn.functional.conv1d(torch.nn.functional.pad(go, (pad, tb.shape[1])), tb.unsqueeze(1), stride=16)
    an = O.reverse_half(an[..., :64])
    assert torch.allclose(an, gyb, atol=1e-12)


@pytest.mark.parametrize("name", ["v2", "v2_small", "v3"])
def test_state_dict_contract_full_size(name):
    """same keys and shapes as the reference's modules at full size."""
    ks = torch.load(os.path.join(GOLDEN, "state_dict_keys.pt"), weights_only=False)[name]
    _, enc, dec = configs.make_autoencoder(name)
    sd = {"encoder." + k: tuple(v.shape) for k, v in enc.state_dict().items()}
    sd.update({"decoder." + k: tuple(v.shape) for k, v in dec.state_dict().items()})
    assert sd == ks


@pytest.mark.parametrize("name", ["v2", "v3", "discrete"])
def test_full_rave_state_dict_contract(name):
    """The whole `rave.RAVE` module tree (pqmf.*, encoder.*, decoder.*, discriminator.*, buffers; dumped from the
    reference built with the gin bindings of configs/{v2,v3,discrete}.gin): same keys, shapes and dtypes."""
    ks = torch.load(os.path.join(GOLDEN, "state_dict_keys.pt"), weights_only=False)["rave_" + name]
    m = configs.build_rave(name)
    sd = {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
    assert set(sd) == set(ks), sorted(set(sd) ^ set(ks))[:20]
    assert sd == ks, [k for k in sd if sd[k] != ks[k]][:20]


def test_tiny_golden_state_dicts_load_strictly():
    for name, kw in (("v2_tiny", {}), ("v2_tiny_causal", dict(padding_mode="causal")),
                     ("v3_tiny", dict(name="v3")), ("v2_small_tiny", dict(ratios=[4, 2, 2, 2]))):
        g = torch.load(os.path.join(GOLDEN, f"autoencoder_{name}.pt"), weights_only=False)
        kw = dict(kw)
        arch = kw.pop("name", "v2")
        pq, enc, dec = configs.make_autoencoder(arch, capacity=8, latent_size=16, **kw)
        holder = nn.Module()
        holder.pqmf, holder.encoder, holder.decoder = pq, enc, dec
        holder.load_state_dict(g["state_dict"], strict=True)
    d = torch.load(os.path.join(GOLDEN, "discriminator_v2.pt"), weights_only=False)
    holder = nn.Module()
    holder.discriminator = configs.make_discriminator_v2(capacity=4)
    holder.load_state_dict(d["state_dict"], strict=True)


def test_remove_weight_norm_hook_is_torch_compatible():
    conv = blocks.normalization(cc.Conv1d(4, 8, 3, padding=(1, 1)))
    assert {"weight_g", "weight_v"} <= set(dict(conv.named_parameters()))
    assert conv.weight_g.shape == (8, 1, 1)
    from torch.nn.utils.weight_norm import WeightNorm
    assert any(isinstance(h, WeightNorm) for h in conv._forward_pre_hooks.values())
    convt = blocks.normalization(cc.ConvTranspose1d(16, 8, 4, stride=2, padding=1))
    assert convt.weight_g.shape == (16, 1, 1)          # dim 0 = Cin for transposed convs


def test_rave_model_builds_and_schedules():
    m = configs.build_rave("v2", capacity=8, latent_size=16, disc_capacity=4, phase_1_duration=2)
    g, d = m.configure_optimizers()
    assert g["optimizer"].defaults["lr"] == 1e-3 and d["optimizer"].defaults["lr"] == 1e-4
    assert g["optimizer"].defaults["betas"] == (.5, .9)
    assert not m.is_discriminator_step(0)
    m.warmed_up = True
    assert m.is_discriminator_step(0) and not m.is_discriminator_step(1) and m.is_discriminator_step(4)
    cb = rave_b200.WarmupCallback()
    m.warmed_up = False
    for i in range(3):
        cb.on_train_batch_start(None, m, None, i)
    assert m.warmed_up
    b = rave_b200.BetaWarmupCallback(initial_value=1e-6, target_value=5e-2, warmup_len=20000)
    b.on_train_batch_start(None, m, None, 0)
    assert 1e-6 < m.beta_factor < 1.1e-6


def test_stacked_l1_terms_match_the_per_feature_loop():
    """core.stacked_l1_terms (the Descript feature taps' sums turned into feature-matching terms by one stacked pass)
    == the reference's per-feature loop: mean over a scale's features of mean|r - f| (/ mean|r|), mean over scales
    (rave/model.py:353-361, rave/core.py:220-252)."""
    import torch
    from rave_b200 import core
    torch.manual_seed(0)
    scales = [[(torch.randn(2, 5, 7), torch.randn(2, 5, 7)) for _ in range(n)] for n in (3, 5, 2)]
    for relative in (False, True):
        leaves, tapped = [], []
        want = 0.
        for feats in scales:
            terms = []
            for r, f in feats:
                r1, f1 = r.clone().requires_grad_(True), f.clone().requires_grad_(True)
                leaves += [r1, f1]
                st = torch.stack([(r1 - f1).abs().sum(), r1.abs().sum()])
                tapped.append((st, r.numel(), 1.0 / (len(feats) * len(scales))))
                terms.append(core.mean_difference(r, f, "L1", relative))
            want = want + sum(terms) / len(terms)
        want = want / len(scales)
        got = core.stacked_l1_terms(tapped, relative)
        assert abs(float(got) - float(want)) < 1e-6 * abs(float(want))
        grads = torch.autograd.grad(got, leaves)
        k = 0
        for feats in scales:
            for r, f in feats:
                r2, f2 = r.clone().requires_grad_(True), f.clone().requires_grad_(True)
                t = core.mean_difference(r2, f2, "L1", relative) / (len(feats) * len(scales))
                gr, gf = torch.autograd.grad(t, [r2, f2])
                assert torch.allclose(grads[k], gr, atol=1e-7) and torch.allclose(grads[k + 1], gf, atol=1e-7)
                k += 2


def test_integration_table_names_every_entry_point():
    """INTEGRATION.md section 3 maps every C entry point of include/rave_b200.h onto the reference call site it replaces."""
    import os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "rave_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(rave_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 60
    missing = []
    for n in names:
        base = re.sub(r"_(fwd|bwd)$", "", n)
        if n not in doc and base not in doc:
            missing.append(n)
    assert not missing, missing


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
    keys, the oracle port as `cpu_baseline`, zero host<->device bytes."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "audio-seconds/s" and line["higher_is_better"] is True
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0
