"""TEST INFRASTRUCTURE -- loads the UNMODIFIED reference (acids-ircam/RAVE) from
/root/reference under third-party stubs, so that it can be executed on CPU in the
build container to (a) pin oracle/rave_oracle.py and (b) generate tests/golden/*.

Nothing here travels to the GPU box in a usable form (/root/reference does not
exist there): only oracle/make_golden.py and the build-container-only tests use
this file.  The product (rave_b200/) never imports it.

Stub recipe = SURVEY.md Appendix C:
  * gin            -> decorators are the identity (bindings are passed as kwargs)
  * cached_conv    -> restatement of the NON-cached classes of cached-conv 2.5.0
                      (call sites rave/pqmf.py:256-273,
                      rave/blocks.py:36,64-76,96-108,...): F.pad + nn.Conv1d
  * pytorch_lightning -> LightningModule = nn.Module
  * GPUtil / librosa / lmdb -> empty modules
  * scipy compat   -> scipy.signal.kaiser, firwin(nyq=) (reference pins scipy 1.10)
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("RAVE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "rave", "pqmf.py"))


# --------------------------------------------------------------------------- gin
def _make_gin():
    gin = types.ModuleType("gin")

    def configurable(*args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]

        def deco(fn):
            return fn

        return deco

    def external_configurable(fn, *a, **k):
        return fn

    def get_configurable(name):
        raise ValueError(name)

    gin.configurable = configurable
    gin.external_configurable = external_configurable
    gin.get_configurable = get_configurable
    gin.add_config_file_search_path = lambda *a, **k: None
    gin.enter_interactive_mode = lambda *a, **k: None
    gin.operative_config_str = lambda *a, **k: ""
    gin.parse_config_files_and_bindings = lambda *a, **k: None
    gin.bind_parameter = lambda *a, **k: None
    gin_torch = types.ModuleType("gin.torch")
    gin.torch = gin_torch
    return gin, gin_torch


# ------------------------------------------------------------------ cached_conv
class _CCState:
    bias = False          # configs/v1.gin:33-34  cc.Conv1d.bias = False
    pad_mode = "centered"  # configs/causal.gin:5  cc.get_padding.mode = 'causal'


def _make_cached_conv():
    cc = types.ModuleType("cached_conv")
    cc.MAX_BATCH_SIZE = 64
    cc.USE_BUFFER_CONV = False
    cc._state = _CCState

    def get_padding(kernel_size, stride=1, dilation=1, mode=None):
        mode = mode if mode is not None else _CCState.pad_mode
        if kernel_size == 1:
            return (0, 0)
        p = (kernel_size - 1) * dilation + 1
        if mode == "centered":
            # even p: the extra sample goes to the right (the one
            # recalled, un-verifiable constant of cached-conv 2.5.0)
            return ((p - 1) // 2, p // 2)
        elif mode == "causal":
            return (p // 2 + (p - 1) // 2, 0)
        raise Exception(f"Padding mode {mode} is not valid")

    class Conv1d(nn.Conv1d):
        def __init__(self, *args, **kwargs):
            self._pad = kwargs.get("padding", (0, 0))
            if isinstance(self._pad, int):
                self._pad = (self._pad, self._pad)
            kwargs["padding"] = 0
            kwargs.pop("cumulative_delay", None)
            if "bias" not in kwargs and len(args) < 8:
                kwargs["bias"] = _CCState.bias
            super().__init__(*args, **kwargs)
            self.cumulative_delay = 0

        def script_cache(self):
            pass

        def forward(self, x):
            x = nn.functional.pad(x, self._pad)
            return nn.functional.conv1d(x, self.weight, self.bias, self.stride,
                                        self.padding, self.dilation, self.groups)

    class ConvTranspose1d(nn.ConvTranspose1d):
        def __init__(self, *args, **kwargs):
            kwargs.pop("cumulative_delay", None)
            if "bias" not in kwargs:
                kwargs["bias"] = _CCState.bias
            super().__init__(*args, **kwargs)
            self.cumulative_delay = 0

    class CachedSequential(nn.Sequential):
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

    class AlignBranches(nn.Module):
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

    def use_cached_conv(state):
        if state:
            raise NotImplementedError("oracle stub only restates the non-cached mode")

    cc.get_padding = get_padding
    cc.Conv1d = Conv1d
    cc.ConvTranspose1d = ConvTranspose1d
    cc.CachedSequential = CachedSequential
    cc.Sequential = CachedSequential
    cc.AlignBranches = AlignBranches
    cc.CachedPadding1d = CachedPadding1d
    cc.use_cached_conv = use_cached_conv
    convs = types.ModuleType("cached_conv.convs")
    convs.get_padding = get_padding
    cc.convs = convs
    return cc, convs


# ----------------------------------------------------------- pytorch_lightning
def _make_pl():
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    pl.Callback = object
    cb = types.ModuleType("pytorch_lightning.callbacks")

    class ModelCheckpoint:
        def __init__(self, *a, **k):
            pass

    cb.ModelCheckpoint = ModelCheckpoint
    cb.Callback = object
    pl.callbacks = cb
    tr = types.ModuleType("pytorch_lightning.trainer")
    st = types.ModuleType("pytorch_lightning.trainer.states")

    class RunningStage:
        SANITY_CHECKING = "sanity_check"

    st.RunningStage = RunningStage
    tr.states = st
    pl.trainer = tr
    return {"pytorch_lightning": pl, "pytorch_lightning.callbacks": cb,
            "pytorch_lightning.trainer": tr, "pytorch_lightning.trainer.states": st}


def _scipy_compat():
    import scipy.signal as ss
    import scipy.signal.windows as win
    if not hasattr(ss, "kaiser"):
        ss.kaiser = win.kaiser
    if not getattr(ss.firwin, "_rave_compat", False):
        orig = ss.firwin

        def firwin(numtaps, cutoff, *args, nyq=None, **kwargs):
            if nyq is not None:
                kwargs["fs"] = 2 * nyq
            return orig(numtaps, cutoff, *args, **kwargs)

        firwin._rave_compat = True
        ss.firwin = firwin


_LOADED = None


def load_reference():
    """Return a namespace with the reference's modules: pqmf, core, blocks,
    discriminator, descript_discriminator, quantization, model, and `cc` (the stub)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    gin, gin_torch = _make_gin()
    sys.modules["gin"] = gin
    sys.modules["gin.torch"] = gin_torch
    cc, convs = _make_cached_conv()
    sys.modules["cached_conv"] = cc
    sys.modules["cached_conv.convs"] = convs
    sys.modules.update(_make_pl())
    for name in ("GPUtil", "librosa", "lmdb"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    _scipy_compat()

    pkg = types.ModuleType("rave")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "rave")]
    sys.modules["rave"] = pkg
    ns = types.SimpleNamespace(cc=cc)
    for name in ("pqmf", "core", "quantization", "blocks", "discriminator",
                 "descript_discriminator", "model"):
        path = os.path.join(REFERENCE_ROOT, "rave", name + ".py")
        spec = importlib.util.spec_from_file_location("rave." + name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules["rave." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
        setattr(ns, name, mod)
    _LOADED = ns
    return ns


def set_padding_mode(mode: str):
    _CCState.pad_mode = mode
