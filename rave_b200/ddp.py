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
        self._flats = {}

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

    def _flat_for(self, bucket: List[torch.nn.Parameter]) -> torch.Tensor:
        """Persistent flat fp32 buffer of one bucket (same parameters -> same buffer on every step, so a captured CUDA
        graph and NCCL's buffer registration see stable addresses)."""
        n = sum(p.numel() for p in bucket)
        key = tuple(id(p) for p in bucket)
        hit = self._flats.get(key)
        if hit is None or hit.numel() != n or hit.device != bucket[0].grad.device:
            hit = self._flats[key] = torch.empty(n, dtype=torch.float32, device=bucket[0].grad.device)
        return hit

    def __call__(self, params: Iterable[torch.nn.Parameter]) -> None:
        """Three passes over the gradients instead of seven: ONE multi-tensor pack into the bucket's persistent flat
        buffer, the all-reduce (AVG inside NCCL: no separate division), and `p.grad` re-pointed at its slice of the
        flat buffer (no copy back; the optimiser reads the averaged gradients in place)."""
        ws = self.world_size(self.group)
        if ws == 1:
            return
        avg = dist.get_backend(self.group) == "nccl"        # gloo (CPU tests) has no AVG: SUM, then one division
        work = []
        for bucket in self.buckets(list(params)):
            flat = self._flat_for(bucket)
            views, off = [], 0
            for p in bucket:
                n = p.numel()
                views.append(flat[off:off + n].view_as(p.grad))
                off += n
            torch._foreach_copy_(views, [p.grad for p in bucket])
            h = dist.all_reduce(flat, op=dist.ReduceOp.AVG if avg else dist.ReduceOp.SUM, group=self.group,
                                async_op=self.async_op)
            work.append((h, flat, bucket, views))
            self.bytes_reduced += flat.numel() * 4
            self.n_collectives += 1
        for h, flat, bucket, views in work:
            if h is not None and self.async_op:
                h.wait()
            if not avg:
                flat.div_(ws)
            for p, v in zip(bucket, views):
                p.grad = v


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (DDP constructor semantics)."""
    if GradientAllReducer.world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


def broadcast_buffers(module: torch.nn.Module, src: int = 0, group=None) -> int:
    """DDP `broadcast_buffers=True` semantics (what pytorch-lightning's DDP strategy gives the reference): before a
    forward every rank takes rank `src`'s floating-point buffers.  The only buffers that move during training are the
    RVQ's EMA state of the `discrete` configuration (`cluster_size`, `embed`, `embed_avg`, `inited`;
    rave/quantization.py:168-179) -- each rank updates them from its own shard, and like the reference under DDP the
    codebooks then FOLLOW RANK 0.  (All-reducing the EMA statistics instead would use every shard, but is a semantic
    change with respect to the reference; not done.)  One flat broadcast; returns the number of elements sent."""
    if GradientAllReducer.world_size(group) == 1:
        return 0
    bufs = [b for b in module.buffers() if b.is_floating_point() and b.numel() > 0]
    if not bufs:
        return 0
    flat = torch.cat([b.detach().reshape(-1).float() for b in bufs])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for b in bufs:
        n = b.numel()
        b.data.copy_(flat[off:off + n].view_as(b))
        off += n
    return int(flat.numel())
