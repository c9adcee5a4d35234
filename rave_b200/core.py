"""Loss helpers -- the subset of rave/core.py the training step calls (SURVEY.md section 8f.1,
a "next" row: cuFFT-backed `torch.stft` + elementwise reductions run on the device through
PyTorch for now; the conv / PQMF hot path they consume is native).
"""
from typing import Callable, Optional, Sequence

import torch
import torch.nn as nn


def mod_sigmoid(x):
    return 2 * torch.sigmoid(x) ** 2.3 + 1e-7


def get_augmented_latent_size(latent_size: int, noise_augmentation: int):
    return latent_size + noise_augmentation


def hinge_gan(score_real, score_fake):
    """rave/core.py:151-155."""
    loss_dis = (torch.relu(1 - score_real) + torch.relu(1 + score_fake)).mean()
    loss_gen = -score_fake.mean()
    return loss_dis, loss_gen


def ls_gan(score_real, score_fake):
    loss_dis = ((score_real - 1).pow(2) + score_fake.pow(2)).mean()
    loss_gen = (score_fake - 1).pow(2).mean()
    return loss_dis, loss_gen


def nonsaturating_gan(score_real, score_fake):
    score_real = torch.clamp(torch.sigmoid(score_real), 1e-7, 1 - 1e-7)
    score_fake = torch.clamp(torch.sigmoid(score_fake), 1e-7, 1 - 1e-7)
    loss_dis = -(torch.log(score_real) + torch.log(1 - score_fake)).mean()
    loss_gen = -torch.log(score_fake).mean()
    return loss_dis, loss_gen


def valid_signal_crop(x, left_rf, right_rf):
    """rave/core.py:220-225."""
    dim = x.shape[1]
    x = x[..., left_rf.item() // dim:]
    if right_rf.item():
        x = x[..., :-right_rf.item() // dim]
    return x


def mean_difference(target, value, norm: str = "L1", relative: bool = False):
    """rave/core.py:236-252.  L1 on CUDA fp32 tensors: both sums from one library pass (ops.l1_stats)."""
    if (norm == "L1" and target.is_cuda and target.dtype == torch.float32 and value.dtype == torch.float32
            and target.shape == value.shape and target.numel() > 0):
        from . import ops
        st = ops.l1_stats(target, value)
        return st[0] / st[1] if relative else st[0] / target.numel()
    diff = target - value
    if norm == "L1":
        diff = diff.abs().mean()
        if relative:
            diff = diff / target.abs().mean()
        return diff
    elif norm == "L2":
        diff = (diff * diff).mean()
        if relative:
            diff = diff / (target * target).mean()
        return diff
    raise Exception(f"Norm must be either L1 or L2, got {norm}")


def mean_difference_halves(base, n_true: int, relative: bool = False):
    """mean_difference(real, fake, 'L1', relative) where real / fake are the two halves (dim 0) of ONE dense fp32 buffer
    whose padding (if any) is zero in both halves; `n_true` = the number of real feature elements (the mean's
    denominator, rave/core.py:244)."""
    from . import ops
    st = ops.l1_halves(base)
    return st[0] / st[1] if relative else st[0] / n_true


_STACKED_WEIGHTS = {}


def stacked_l1_terms(tapped, relative: bool = False):
    """sum_i w_i * mean_difference_i for features whose (sum |real - fake|, sum |real|) pairs already exist:
    `tapped` = [(sums[2], n_true, w)].  One stack, one multiply (or divide + multiply), one sum -- instead of a scalar
    division and an addition (and their backward launches) per feature.  The constant weight vector lives on the device,
    built once per (device, weights) outside any stream capture."""
    S = torch.stack([t[0] for t in tapped])                                  # [n, 2]
    key = (str(S.device), bool(relative), tuple((int(t[1]), float(t[2])) for t in tapped))
    w = _STACKED_WEIGHTS.get(key)
    if w is None:
        if S.is_cuda and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("stacked_l1_terms: first call for these features inside a stream capture (run one eager "
                               "step first: the weight vector is uploaded once)")
        vals = [t[2] if relative else t[2] / t[1] for t in tapped]
        w = _STACKED_WEIGHTS[key] = torch.tensor(vals, dtype=torch.float64).to(S.dtype).to(S.device)
    if relative:
        return ((S[:, 0] / S[:, 1]) * w).sum()
    return (S[:, 0] * w).sum()


class _StftWindow(nn.Module):
    """Holder of one scale's hann window (`stfts.<i>.window`, the key torchaudio.transforms.Spectrogram contributes)."""

    def __init__(self, n_fft: int) -> None:
        super().__init__()
        self.n_fft = n_fft
        self.register_buffer("window", torch.hann_window(n_fft))


class MultiScaleSTFT(nn.Module):
    """rave/core.py:269-319 (magnitude spectrogram, hann window, hop = n_fft/4, centred)."""

    def __init__(self, scales: Sequence[int], sample_rate: int, magnitude: bool = True,
                 normalized: bool = False, num_mels: Optional[int] = None) -> None:
        super().__init__()
        if num_mels is not None:
            raise NotImplementedError("mel scales need librosa, absent here; not used by v2/v3/discrete")
        self.scales = scales
        self.magnitude = magnitude
        self.normalized = normalized
        # the reference keeps one torchaudio Spectrogram per scale in `self.stfts` (rave/core.py:283-296), whose hann
        # `window` buffers are part of RAVE.state_dict(): same names here
        self.stfts = nn.ModuleList([_StftWindow(s) for s in scales])
        for s in scales:
            bw = torch.full((s // 2 + 1,), 0.5 * s)       # rfft backward as one c2r transform (ops.RfftFn)
            bw[0] = s
            bw[-1] = s
            self.register_buffer(f"rfft_bw_{s}", bw, persistent=False)

    def __getattr__(self, name):
        if name.startswith("window_"):                  # window_<n_fft>: the persistent buffer stfts[i].window
            s = int(name[7:])
            return self.stfts[list(self.scales).index(s)].window
        return super().__getattr__(name)

    def complex_stfts(self, x):
        x = x.reshape(-1, x.shape[-1])
        if x.is_cuda and x.dtype == torch.float32 and not self.normalized:
            # library framing kernel (pad + frame + window; adjoint = window + overlap-add + fold) and cuFFT for
            # the transform; like torch.stft the result is a [N, bins, frames] view of a [N, frames, bins] buffer
            from . import ops
            if all(x.shape[-1] > s // 2 for s in self.scales):
                return [ops.rfft(ops.stft_frames(x, getattr(self, f"window_{s}"), s, s // 4),
                                 getattr(self, f"rfft_bw_{s}")).transpose(-1, -2) for s in self.scales]
        return [torch.stft(x, s, hop_length=s // 4, win_length=s, window=getattr(self, f"window_{s}"),
                           center=True, pad_mode="reflect", normalized=self.normalized, onesided=True,
                           return_complex=True) for s in self.scales]

    def forward(self, x):
        return [y.abs() if self.magnitude else torch.stack([y.real, y.imag], -1) for y in self.complex_stfts(x)]


class AudioDistanceV1(nn.Module):
    """rave/core.py:322-344."""

    def __init__(self, multiscale_stft: Callable[[], nn.Module], log_epsilon: float) -> None:
        super().__init__()
        self.multiscale_stft = multiscale_stft()
        self.log_epsilon = log_epsilon

    def forward(self, x, y):
        mstft = self.multiscale_stft
        if (x.is_cuda and not x.requires_grad and isinstance(mstft, MultiScaleSTFT) and mstft.magnitude
                and x.dtype == torch.float32):
            # fused path: one kernel per scale for the whole |.|, log, L2-relative + L1 tail (and one for its
            # gradient) instead of ~40 ATen launches; cuFFT still does the transforms
            from . import ops
            distance = 0.
            for sx, sy in zip(mstft.complex_stfts(x), mstft.complex_stfts(y)):
                distance = distance + ops.spectral_distance(sx, sy, self.log_epsilon)
            return {"spectral_distance": distance}
        stfts_x = self.multiscale_stft(x)
        stfts_y = self.multiscale_stft(y)
        distance = 0.
        for sx, sy in zip(stfts_x, stfts_y):
            logx = torch.log(sx + self.log_epsilon)
            logy = torch.log(sy + self.log_epsilon)
            distance = distance + mean_difference(sx, sy, norm="L2", relative=True) \
                + mean_difference(logx, logy, norm="L1")
        return {"spectral_distance": distance}


@torch.enable_grad()
def get_rave_receptive_field(model, n_channels=1):
    """rave/core.py:180-217: autograd probe of the input gradient's support."""
    N = 2 ** 15
    model.eval()
    device = next(iter(model.parameters())).device
    while True:
        x = torch.randn(1, model.n_channels, N, requires_grad=True, device=device)
        z = model.encode(x)
        z = model.encoder.reparametrize(z)[0]
        y = model.decode(z)
        y[0, 0, N // 2].backward()
        grad = x.grad.data.reshape(-1)
        left_grad, right_grad = grad.chunk(2, 0)
        if (left_grad[0] == 0) and right_grad[-1] == 0:
            break
        N *= 2
    left_rf = len(left_grad[left_grad != 0])
    right_rf = len(right_grad[right_grad != 0])
    model.zero_grad()
    return left_rf, right_rf


# ---------------------------------------------------------------------------------------------
# FFT noise filtering helpers of NoiseGeneratorV2 (rave/core.py:48-81) -- cuFFT through torch,
# SURVEY row 8f.4 ("next"); the strided convs that feed them are on the library kernels.
# ---------------------------------------------------------------------------------------------

def amp_to_impulse_response(amp, target_size):
    """Zero-phase band amplitudes -> windowed, causal-shifted FIR of length `target_size`."""
    spec = torch.complex(amp, torch.zeros_like(amp))
    ir = torch.fft.irfft(spec)
    n = ir.shape[-1]
    ir = torch.roll(ir, n // 2, -1) * torch.hann_window(n, dtype=ir.dtype, device=ir.device)
    ir = nn.functional.pad(ir, (0, int(target_size) - int(n)))
    return torch.roll(ir, -n // 2, -1)


def fft_convolve(signal, kernel):
    """Linear convolution on the last axis via zero-padded rFFT, keeping the last half."""
    signal = nn.functional.pad(signal, (0, signal.shape[-1]))
    kernel = nn.functional.pad(kernel, (kernel.shape[-1], 0))
    out = torch.fft.irfft(torch.fft.rfft(signal) * torch.fft.rfft(kernel))
    return out[..., out.shape[-1] // 2:]
