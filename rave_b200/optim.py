"""Adam on the library's multi-tensor kernel (one launch per <= 96 parameter tensors).

Same arithmetic and state layout as `torch.optim.Adam(params, lr, betas)` without weight decay / amsgrad
(reference: rave/model.py:226-236): per-parameter `exp_avg` / `exp_avg_sq`, one shared step counter.  `lr` and the
step counter are device tensors, so a captured CUDA graph replays the update and `LinearLR` (which `fill_`s a tensor
lr in place) keeps working.  CUDA fp32 parameters only: there is no CPU path."""
import ctypes
from typing import Iterable, Tuple

import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8):
        params = list(params)
        if not params or not all(p.is_cuda and p.dtype == torch.float32 for p in params):
            raise _lib.RaveB200Error("FusedAdam needs CUDA fp32 parameters (there is no CPU path)")
        dev = params[0].device
        if not torch.is_tensor(lr):
            lr = torch.tensor(float(lr), dtype=torch.float32, device=dev)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        for group in self.param_groups:
            group["step"] = torch.zeros((), dtype=torch.float32, device=dev)

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            n = len(ps)
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            states = [self._state(p) for p in ps]
            for p, g in zip(ps, grads):
                if not p.is_contiguous() or g.dtype != torch.float32:
                    raise _lib.RaveB200Error("FusedAdam: parameters must be contiguous with fp32 gradients")
            Arr = ctypes.c_void_p * n
            pa = Arr(*[p.data_ptr() for p in ps])
            ga = Arr(*[g.data_ptr() for g in grads])
            ma = Arr(*[s["exp_avg"].data_ptr() for s in states])
            va = Arr(*[s["exp_avg_sq"].data_ptr() for s in states])
            na = (ctypes.c_long * n)(*[p.numel() for p in ps])
            lr = group["lr"]
            if not (torch.is_tensor(lr) and lr.is_cuda and lr.dtype == torch.float32):
                raise _lib.RaveB200Error("FusedAdam: lr must stay a CUDA fp32 tensor")
            b1, b2 = group["betas"]
            _lib.call("rave_adam_multi", n, pa, ga, ma, va, na, lr.data_ptr(), group["step"].data_ptr(), float(b1),
                      float(b2), float(group["eps"]), _lib.stream_ptr())
        # The kernel writes the parameters through raw pointers: autograd's version counters do not move, so the
        # engine's per-module cache of prepared (weight-normalised, tap-major bf16) weights would keep serving the
        # pre-update values.  Drop it (a host-side epoch counter; captured graphs never use the cache).
        from . import engine
        engine.invalidate_prepared()
        return loss
