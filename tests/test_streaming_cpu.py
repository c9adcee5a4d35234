"""CPU: the HOST logic of the streaming (cached) convolutions -- ring buffers, stride-alignment delays, the transposed
conv's overlap-add tail, AlignBranches delay lines, cumulative_delay bookkeeping -- with the CUDA library ops replaced
by torch stand-ins (test infrastructure; the product has no CPU path).  Property checked: what the reference's
tests/test_residual.py checks for cached_conv -- chunked streaming output == offline output delayed by
`cumulative_delay`.  The GPU twin (same property through the real kernels) is
tests/test_gpu_parity.py::test_streaming_cached_convs_reproduce_offline."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.conftest import rel_l2


def _act(x, act, slope, alpha):
    from rave_b200 import ops
    if act == ops.ACT_NONE:
        return x
    if act == ops.ACT_LEAKY:
        return F.leaky_relu(x, slope)
    a = alpha.reshape(1, -1, 1)
    return x + torch.sin(a * x) ** 2 / (a + 1e-9)


@pytest.fixture
def torch_ops(monkeypatch):
    from rave_b200 import ops

    def conv1d(x, w, bias=None, res=None, stride=1, dilation=1, pad=(0, 0), act=ops.ACT_NONE, slope=0.2, alpha=None):
        y = F.conv1d(F.pad(_act(x, act, slope, alpha), pad), w, bias, stride, 0, dilation,
                     groups=x.shape[1] // w.shape[1])
        return y if res is None else y + res

    def conv_transpose1d(x, w, bias=None, stride=1, padding=0, act=ops.ACT_NONE, slope=0.2, alpha=None):
        return F.conv_transpose1d(_act(x, act, slope, alpha), w, bias, stride, padding)

    def weight_norm(v, g):
        return g * v / v.flatten(1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))

    monkeypatch.setattr(ops, "conv1d", conv1d)
    monkeypatch.setattr(ops, "conv_transpose1d", conv_transpose1d)
    monkeypatch.setattr(ops, "weight_norm", weight_norm)
    monkeypatch.setattr(ops, "activation", lambda x, act, slope=0.2, alpha=None: _act(x, act, slope, alpha))


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
    with torch.no_grad():
        y_off = off(x)
        y_on = torch.cat([on(c) for c in x.split(chunk, -1)], -1)
    assert y_on.shape == y_off.shape
    return y_on, y_off, on.cumulative_delay


def _one(mod, mode="centered"):
    from rave_b200 import cc

    def build():
        with cc.configure(conv_bias=True, padding_mode=mode):
            return mod()
    return build


def test_single_modules_centred_padding_delay(torch_ops):
    from rave_b200 import blocks, cc
    x = torch.randn(2, 16, 1024)
    cases = [
        (lambda: blocks.Residual(blocks.DilatedUnit(16, 3, 3)), 3),
        (lambda: blocks.normalization(cc.Conv1d(16, 32, 8, stride=4, padding=cc.get_padding(8, 4))), 1),
        (lambda: blocks.normalization(cc.ConvTranspose1d(16, 8, 8, stride=4, padding=2)), 2),
        (lambda: blocks.normalization(cc.Conv1d(16, 16, 7, padding=cc.get_padding(7))), 3),
        (lambda: blocks.normalization(cc.Conv1d(16, 16, 3, dilation=9, padding=cc.get_padding(3, dilation=9))), 9),
    ]
    for mod, want_d in cases:
        for chunk in (64, 128):
            y_on, y_off, d = _stream_vs_offline(_one(mod), x, chunk)
            assert d == want_d, (d, want_d)
            assert rel_l2(y_on[..., d:], y_off[..., :y_off.shape[-1] - d]) < 1e-5, (want_d, chunk)


def test_causal_stack_streams_with_the_transposed_convs_delay(torch_ops):
    """Encoder / decoder style stack with causal padding (configs/causal.gin): zero delay from the convs, 2 samples from
    the transposed conv's symmetric crop; outputs after the start-up transient are identical to offline."""
    from rave_b200 import blocks, cc

    def build():
        with cc.configure(conv_bias=True, padding_mode="causal"):
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
    x = torch.randn(2, 16, 2048)
    for chunk in (128, 256, 1024):
        y_on, y_off, d = _stream_vs_offline(build, x, chunk)
        warm = 16
        assert rel_l2(y_on[..., 2 + warm:], y_off[..., warm:-2]) < 1e-5, chunk
        # and the transient really is confined to the first samples
        assert rel_l2(y_on[..., 2:2 + warm], y_off[..., :warm]) < 5e-2


def test_stream_is_chunk_size_invariant(torch_ops):
    """The streamed output does not depend on how the signal is cut (the cache carries exactly the missing context)."""
    from rave_b200 import blocks, cc

    def build():
        with cc.configure(conv_bias=True, padding_mode="centered"):
            return cc.CachedSequential(
                blocks.normalization(cc.Conv1d(8, 16, 7, padding=cc.get_padding(7))),
                blocks.Residual(blocks.DilatedUnit(16, 3, 3)),
                nn.LeakyReLU(.2),
                blocks.normalization(cc.Conv1d(16, 16, 8, stride=4, padding=cc.get_padding(8, 4))),
                nn.LeakyReLU(.2),
                blocks.normalization(cc.ConvTranspose1d(16, 8, 8, stride=4, padding=2)))
    x = torch.randn(1, 8, 1024)
    outs = [_stream_vs_offline(build, x, c)[0] for c in (64, 256, 1024)]
    assert rel_l2(outs[0], outs[1]) < 1e-6 and rel_l2(outs[0], outs[2]) < 1e-6


def test_streaming_cached_pqmf_reproduces_offline_operator(torch_ops):
    """CachedPQMF built under cc.use_cached_conv(True) (rave/pqmf.py:245-294 in streaming mode): analysis and synthesis fed
    chunk by chunk reproduce the offline operator (golden-pinned: tests/golden/pqmf.pt; here restated with torch convs
    on the module's own taps) delayed by the convs' cumulative delays."""
    from rave_b200 import cc
    from rave_b200.pqmf import CachedPQMF, reverse_half
    cc.use_cached_conv(True)
    try:
        pq = CachedPQMF(100, 16)
    finally:
        cc.use_cached_conv(False)
    assert pq.streaming
    x = torch.randn(2, 1, 16 * 512)
    # offline restatement on the same taps: rave/pqmf.py:279-294 with zero-padded (non-cached) convs
    wf, wi = pq.forward_conv.weight, pq.inverse_conv.weight
    pf = (wf.shape[-1] - 1) // 2
    mb_off = reverse_half(F.conv1d(x, wf, None, 16, pf))
    pi = (wi.shape[-1] - 1) // 2
    y = F.conv1d(reverse_half(mb_off), wi, None, 1, pi) * 16
    y = y.flip(1).permute(0, 2, 1)
    y = y.reshape(y.shape[0], y.shape[1], -1, 16).permute(0, 2, 1, 3)
    y_off = y.reshape(y.shape[0], y.shape[1], -1)
    with torch.no_grad():
        mb_on = torch.cat([pq(c) for c in x.split(2048, -1)], -1)
        d_f = pq.forward_conv.cumulative_delay            # in multiband samples
        assert mb_on.shape == mb_off.shape
        assert rel_l2(mb_on[..., d_f:], mb_off[..., :mb_off.shape[-1] - d_f]) < 1e-5
        pq2 = pq                                            # synthesis on the OFFLINE bands: isolates the inverse's delay
        y_on = torch.cat([pq2.inverse(c) for c in mb_off.split(128, -1)], -1)
    d_i = pq.inverse_conv.cumulative_delay * 16             # band samples -> audio samples
    assert y_on.shape == y_off.shape
    assert rel_l2(y_on[..., d_i:], y_off[..., :y_off.shape[-1] - d_i]) < 1e-5
    # and the pair is a (delayed) near-perfect reconstruction of the input
    total = d_f * 16 + d_i
    with torch.no_grad():
        cc.use_cached_conv(True)
        try:
            pq3 = CachedPQMF(100, 16)
        finally:
            cc.use_cached_conv(False)
        rec = torch.cat([pq3.inverse(pq3(c)) for c in x.split(2048, -1)], -1)
    lo, hi = total + 1024, x.shape[-1] - 1024
    best = min(rel_l2(rec[..., lo + s:hi + s], x[..., lo - total:hi - total]) for s in (-16, -1, 0, 1, 16))
    assert best < 2e-2, best
