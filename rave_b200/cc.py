"""Host-side mirror of the NON-cached operator layer the reference reaches through the
third-party package `cached-conv>=2.5.0` (requirements.txt:14; semantics: SURVEY.md App. A).

Same names / constructor arguments / attributes as `cached_conv` so the block definitions read
like the reference's (`cc.Conv1d`, `cc.ConvTranspose1d`, `cc.get_padding`, `cc.CachedSequential`,
`cc.AlignBranches`, `cc.CachedPadding1d`, `cc.MAX_BATCH_SIZE`, `cc.USE_BUFFER_CONV`,
`cc.use_cached_conv`), but `forward` launches the sm_100a kernels of librave_b200.so.

gin is not available: the two gin-configurable knobs the reference's configs bind
(`cc.Conv1d.bias = False`, configs/v1.gin:33-34; `cc.get_padding.mode = 'causal'`,
configs/causal.gin:5) are the fields of `cc.config`.
"""
import contextlib
from typing import Tuple

import torch
import torch.nn as nn

from . import ops

MAX_BATCH_SIZE = 64
USE_BUFFER_CONV = False


class _Config:
    conv_bias = False          # cc.Conv1d.bias / cc.ConvTranspose1d.bias
    padding_mode = "centered"  # cc.get_padding.mode


config = _Config()


@contextlib.contextmanager
def configure(conv_bias=None, padding_mode=None):
    old = (config.conv_bias, config.padding_mode)
    if conv_bias is not None:
        config.conv_bias = conv_bias
    if padding_mode is not None:
        config.padding_mode = padding_mode
    try:
        yield
    finally:
        config.conv_bias, config.padding_mode = old


def use_cached_conv(state: bool):
    if state:
        raise NotImplementedError(
            "streaming (cached) convolutions are out of scope (SURVEY.md 8f.4); training uses the "
            "non-cached mode (scripts/train.py never enables it)")


def get_padding(kernel_size: int, stride: int = 1, dilation: int = 1, mode: str = None) -> Tuple[int, int]:
    """'same' padding as a (left, right) pair; `stride` is accepted but unused (App. A)."""
    mode = mode if mode is not None else config.padding_mode
    if kernel_size == 1:
        return (0, 0)
    p = (kernel_size - 1) * dilation + 1
    if mode == "centered":
        return ((p - 1) // 2, p // 2)
    if mode == "causal":
        return (p // 2 + (p - 1) // 2, 0)
    raise Exception(f"Padding mode {mode} is not valid")


def _act_code(act_module):
    """(code, slope, alpha) of an `activation(dim)` module, or None if it cannot be fused."""
    from .blocks import Snake  # local import: blocks imports cc
    if act_module is None:
        return (ops.ACT_NONE, 0.0, None)
    if isinstance(act_module, nn.LeakyReLU):
        return (ops.ACT_LEAKY, float(act_module.negative_slope), None)
    if isinstance(act_module, Snake):
        return (ops.ACT_SNAKE, 0.0, act_module.alpha)
    return None


class Conv1d(nn.Conv1d):
    """cc.Conv1d: explicit asymmetric padding tuple + nn.Conv1d parameters.  `forward(x, act=,
    res=)` optionally fuses the preceding activation module and a residual add into the kernel."""

    def __init__(self, *args, **kwargs):
        pad = kwargs.get("padding", (0, 0))
        if isinstance(pad, int):
            pad = (pad, pad)
        self._pad = tuple(pad)
        kwargs["padding"] = 0
        kwargs.pop("cumulative_delay", None)
        if "bias" not in kwargs and len(args) < 8:
            kwargs["bias"] = config.conv_bias
        super().__init__(*args, **kwargs)
        self.cumulative_delay = 0

    def script_cache(self):
        pass

    def forward(self, x, act=None, res=None):
        code = _act_code(act)
        if code is None:
            x = act(x)
            code = (ops.ACT_NONE, 0.0, None)
        alpha = code[2].reshape(-1) if code[2] is not None else None
        if self.groups != 1:
            # grouped conv (the v1 Encoder's last layer, rave/blocks.py:489-497: groups = n_out): one launch per group on
            # its slice of the input / weight channels, same nn.Conv1d parameter layout [Cout, Cin / groups, K]
            g = self.groups
            cin, cout = self.in_channels // g, self.out_channels // g
            outs = []
            for i in range(g):
                xi = x[:, i * cin:(i + 1) * cin].contiguous()
                ai = alpha[i * cin:(i + 1) * cin].contiguous() if alpha is not None else None
                bi = self.bias[i * cout:(i + 1) * cout] if self.bias is not None else None
                ri = res[:, i * cout:(i + 1) * cout].contiguous() if res is not None else None
                outs.append(ops.conv1d(xi, self.weight[i * cout:(i + 1) * cout], bi, ri, self.stride[0],
                                       self.dilation[0], self._pad, code[0], code[1], ai))
            return torch.cat(outs, 1)
        return ops.conv1d(x, self.weight, self.bias, res, self.stride[0], self.dilation[0], self._pad,
                          code[0], code[1], alpha)


class ConvTranspose1d(nn.ConvTranspose1d):
    def __init__(self, *args, **kwargs):
        kwargs.pop("cumulative_delay", None)
        if "bias" not in kwargs and len(args) < 8:
            kwargs["bias"] = config.conv_bias
        super().__init__(*args, **kwargs)
        if self.groups != 1 or self.output_padding[0] != 0 or self.dilation[0] != 1:
            raise NotImplementedError("only plain ConvTranspose1d is on the hot path")
        self.cumulative_delay = 0

    def script_cache(self):
        pass

    def forward(self, x, act=None):
        code = _act_code(act)
        if code is None:
            x = act(x)
            code = (ops.ACT_NONE, 0.0, None)
        alpha = code[2].reshape(-1) if code[2] is not None else None
        return ops.conv_transpose1d(x, self.weight, self.bias, self.stride[0], self.padding[0], code[0],
                                    code[1], alpha)


def _is_activation(m) -> bool:
    from .blocks import Snake
    return isinstance(m, (nn.LeakyReLU, Snake))


class CachedSequential(nn.Sequential):
    """cc.CachedSequential.  forward fuses every `activation -> conv` pair into one launch."""

    def __init__(self, *args, **kwargs):
        cumulative_delay = kwargs.pop("cumulative_delay", 0)
        stride = kwargs.pop("stride", 1)
        super().__init__(*args, **kwargs)
        last = 0
        for m in reversed(list(self)):
            if hasattr(m, "cumulative_delay"):
                last = m.cumulative_delay
                break
        self.cumulative_delay = cumulative_delay * stride + last

    def _tc_plan(self):
        """Tensor-core plan of this sequence (None if a member is unsupported), cached per mode."""
        from . import engine
        key = self.training
        cache = self.__dict__.setdefault("_tc_plan_cache", {})
        if key not in cache:
            specs = engine.plan_sequential(list(self))
            if specs is not None and not engine.chain_supported(specs):
                specs = None
            cache[key] = specs
        return cache[key]

    def forward(self, x, res=None):
        from . import engine
        mode = engine.precision()
        x3 = mode == "bf16x3"
        if x3 and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            mode = "fp32"          # the split-operand mode is a forward path: gradients run on the fp32 kernels
        if res is None and mode in ("bf16", "bf16x3") and x.is_cuda and x.dim() == 3:
            specs = self._tc_plan()
            if specs is not None and (specs[0].kind != "conv" or x.shape[-1] % specs[0].stride == 0):
                (out,) = engine.run_chain(engine.to_channel_last(x, x3=x3), specs, x3=x3)
                Lout = engine.chain_lengths(specs, x.shape[-1])[-1]
                if out.shape[1] != Lout:
                    out = out[:, :Lout].contiguous()
                return engine.from_channel_last(out)
        mods = list(self)
        i = 0
        n = len(mods)
        while i < n:
            m = mods[i]
            last = i == n - 1
            if _is_activation(m) and i + 1 < n and isinstance(mods[i + 1], (Conv1d, ConvTranspose1d)):
                conv = mods[i + 1]
                if isinstance(conv, Conv1d):
                    x = conv(x, act=m, res=res if i + 1 == n - 1 else None)
                    if i + 1 == n - 1:
                        res = None
                else:
                    x = conv(x, act=m)
                i += 2
                continue
            if isinstance(m, Conv1d) and last and res is not None:
                x = m(x, res=res)
                res = None
            else:
                x = m(x)
            i += 1
        if res is not None:
            x = x + res
        return x


Sequential = CachedSequential


class AlignBranches(nn.Module):
    """All delays are 0 outside the streaming mode: a plain fan-out."""

    def __init__(self, *branches, delays=None, cumulative_delay=0, stride=1):
        super().__init__()
        self.branches = nn.ModuleList(branches)
        self.cumulative_delay = cumulative_delay

    def forward(self, x):
        return [b(x) for b in self.branches]


class CachedPadding1d(nn.Module):
    def __init__(self, padding, crop=False):
        super().__init__()
        self.padding = padding

    def forward(self, x):
        return x
