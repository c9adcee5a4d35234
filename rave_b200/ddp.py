"""Data-parallel gradient exchange: one process per GPU, batch axis sharded, one all-reduce (SUM / N)
of the gradients of the parameter group being stepped.

The reference has no distributed code of its own; multi-GPU training goes through
pytorch-lightning's implicit DDP (scripts/train.py:215-221,242-255), which all-reduces every
parameter's gradient after each backward.  Here the exchange is explicit and restricted to the
group that is about to be stepped (G-step: encoder+generator, D-step: discriminator), in
flat fp32 buckets sized for launch latency rather than link count (NVSwitch: every peer at full
bandwidth).  Works with NCCL (GPU) and gloo (CPU tests).
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientAllReducer:
    """Callable usable as `grad_hook` of `RAVE.training_step`."""

    def __init__(self, bucket_bytes: int = 64 << 20, group: Optional[dist.ProcessGroup] = None,
                 async_op: bool = True):
        self.bucket_elems = max(1, bucket_bytes // 4)
        self.group = group
        self.async_op = async_op
        self.bytes_reduced = 0
        self.n_collectives = 0

    @staticmethod
    def world_size(group=None) -> int:
        return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def buckets(self, params: Iterable[torch.nn.Parameter]) -> List[List[torch.nn.Parameter]]:
        """Greedy packing, in reverse registration order (the order gradients become ready)."""
        out, cur, n = [], [], 0
        for p in reversed([p for p in params if p.grad is not None]):
            if cur and n + p.numel() > self.bucket_elems:
                out.append(cur)
                cur, n = [], 0
            cur.append(p)
            n += p.numel()
        if cur:
            out.append(cur)
        return out

    def __call__(self, params: Iterable[torch.nn.Parameter]) -> None:
        ws = self.world_size(self.group)
        if ws == 1:
            return
        work = []
        for bucket in self.buckets(list(params)):
            flat = torch.cat([p.grad.reshape(-1) for p in bucket])
            h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=self.async_op)
            work.append((h, flat, bucket))
            self.bytes_reduced += flat.numel() * 4
            self.n_collectives += 1
        for h, flat, bucket in work:
            if h is not None and self.async_op:
                h.wait()
            flat.div_(ws)
            off = 0
            for p in bucket:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (DDP constructor semantics)."""
    if GradientAllReducer.world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
