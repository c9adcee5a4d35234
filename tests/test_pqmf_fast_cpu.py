"""CPU: the factorised PQMF tables (rave_b200/pqmf.py::_factorise) and the two-stage evaluation the fast kernels of
csrc/pqmf.cu implement, emulated in torch with the kernels' index conventions, against the oracle's dense operators
(CachedPQMF.forward / inverse, rave/pqmf.py:279-294) and their autograd adjoints."""
import os

import pytest
import torch

from oracle import rave_oracle as O
from rave_b200 import cc, pqmf
from tests.conftest import GOLDEN, rel_l2


def analysis_fast_emul(x, Ct, Qt, Lout, pad_l, flip=True):
    """y[b][k][n] = sgn(k,n) sum_r Ct[r][k] sum_i Qt[i][r] x[b][16 n + 32 i + r - pad_l]"""
    B, T = x.shape
    need = 16 * (Lout - 1) + 32 * 16 + 31 + 1
    xp = torch.zeros(B, pad_l + max(T, need) + 64, dtype=x.dtype)
    xp[:, pad_l:pad_l + T] = x
    n = torch.arange(Lout).view(-1, 1, 1)
    i = torch.arange(17).view(1, -1, 1)
    r = torch.arange(32).view(1, 1, -1)
    idx = 16 * n + 32 * i + r                                        # relative to -pad_l
    g = xp[:, idx]                                                   # [B, Lout, 17, 32]
    u = (g * Qt.view(1, 1, 17, 32)).sum(2)                           # [B, Lout, 32]
    y = torch.einsum("bnr,rk->bkn", u, Ct)
    if flip:
        y[:, 1::2, ::2] *= -1
    return y


def synthesis_fast_emul(x, Cc, Qt, pad_l, scale, flip=True):
    """out[b][16 t + 15 - m] = scale sum_{e<2} sum_{i<17-e} Qt[i][16e+m] v[16e+m][t + 2i + e - pad_l],
    v[r][tau] = sum_c Cc[c][r] sgn(c,tau) x[b][c][tau]"""
    B, M, L = x.shape
    xs = x.clone()
    if flip:
        xs[:, 1::2, ::2] *= -1
    v = torch.einsum("bct,cr->brt", xs, Cc)                          # [B, 32, L]
    vp = torch.zeros(B, 32, pad_l + L + 40, dtype=x.dtype)
    vp[:, :, pad_l:pad_l + L] = v
    out = torch.zeros(B, 16 * L, dtype=x.dtype)
    t = torch.arange(L)
    for m in range(16):
        acc = torch.zeros(B, L, dtype=x.dtype)
        for e in range(2):
            for i in range(17 - e):
                acc += Qt[i, 16 * e + m] * vp[:, 16 * e + m, t + 2 * i + e]
        out[:, 16 * t + 15 - m] = scale * acc
    return out


@pytest.mark.parametrize("mode", ["centered", "causal"])
def test_factorised_tables_reproduce_dense_operators(mode):
    with cc.configure(padding_mode=mode):
        p = pqmf.CachedPQMF(attenuation=100, n_band=16)
    t = p._tables()
    assert isinstance(t["taps"], tuple) and isinstance(t["w"], tuple), "bank not recognised as cosine-modulated"
    hk = p.hk
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 1, 2048, generator=g, dtype=torch.float64)
    Ct, Qt, ntaps = t["taps"]
    Lout = (2048 + t["pad_l"] + t["pad_r"] - ntaps) // 16 + 1
    y = analysis_fast_emul(x[:, 0], Ct.double(), Qt.double(), Lout, t["pad_l"])
    y_ref = O.pqmf_analysis(x.float(), hk, mode)
    assert y.shape == y_ref.shape
    assert rel_l2(y, y_ref) < 1e-6
    yb = torch.randn(2, 16, 128, generator=g, dtype=torch.float64)
    Cc, Qs = t["w"]
    out = synthesis_fast_emul(yb, Cc.double(), Qs.double(), t["w_pad"], 16.0)
    out_ref = O.pqmf_synthesis(yb.float(), hk, mode)
    assert rel_l2(out.view_as(out_ref), out_ref) < 1e-6
    # adjoints: backward of analysis = synthesis form with taps_bwd; backward of synthesis = analysis form with w_bwd
    xo = x.float().clone().requires_grad_(True)
    yo = O.pqmf_analysis(xo, hk, mode)
    gy = torch.randn(yo.shape, generator=g)
    (gx_ref,) = torch.autograd.grad(yo, xo, gy)
    Cb, Qb = t["taps_bwd"]
    gx = synthesis_fast_emul(gy.double(), Cb.double(), Qb.double(), t["taps_bwd_pad"], 1.0)
    assert rel_l2(gx.view_as(gx_ref), gx_ref) < 1e-6
    ybo = yb.float().clone().requires_grad_(True)
    oo = O.pqmf_synthesis(ybo, hk, mode)
    go = torch.randn(oo.shape, generator=g)
    (gy_ref,) = torch.autograd.grad(oo, ybo, go)
    Cwt, Qw, _ = t["w_bwd"]
    gyb = analysis_fast_emul(go[:, 0].double(), Cwt.double(), Qw.double(), 128, t["w_bwd_pad"])
    assert rel_l2(gyb, gy_ref) < 1e-6


def test_factorise_rejects_a_bank_that_is_not_cosine_modulated():
    torch.manual_seed(0)
    assert pqmf._factorise(torch.randn(16, 513)) is None
