"""Residual vector quantiser of the `discrete` configuration (module surface of rave/quantization.py:
EuclideanCodebook / VectorQuantization / ResidualVectorQuantization, same buffer names).

SURVEY K18: N = B*T/1024 rows x 1024 codes x 16 quantisers -- negligible FLOPs next to the conv stacks,
so the arithmetic (one [N,128]x[128,1024] distance GEMM, argmax, EMA scatter) stays on torch/cuBLAS.
Semantics follow the reference (EMA codebooks with Laplace smoothing, k-means initialisation on the
first training batch, dead-code replacement, straight-through estimator, commitment loss).
"""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _pick_rows(rows: torch.Tensor, count: int) -> torch.Tensor:
    """`count` rows drawn from `rows` (without replacement when there are enough)."""
    n = rows.shape[0]
    if n >= count:
        idx = torch.randperm(n, device=rows.device)[:count]
    else:
        idx = torch.randint(0, n, (count,), device=rows.device)
    return rows[idx]


def _nearest_code(rows: torch.Tensor, codes: torch.Tensor) -> torch.Tensor:
    """argmin_j ||rows_i - codes_j||^2 through the expanded form (one GEMM)."""
    ct = codes.t()
    neg_d2 = -(rows.pow(2).sum(1, keepdim=True) - 2 * rows @ ct + ct.pow(2).sum(0, keepdim=True))
    return neg_d2.max(dim=-1).indices


def kmeans(samples: torch.Tensor, num_clusters: int, num_iters: int = 10):
    """Lloyd iterations from a random subset; empty clusters keep their previous centre."""
    dim = samples.shape[-1]
    means = _pick_rows(samples, num_clusters)
    counts = None
    for _ in range(num_iters):
        d2 = ((samples[:, None] - means[None]) ** 2).sum(-1)
        assign = (-d2).max(dim=-1).indices
        counts = torch.bincount(assign, minlength=num_clusters)
        empty = counts == 0
        sums = torch.zeros(num_clusters, dim, dtype=samples.dtype, device=samples.device)
        sums.index_add_(0, assign, samples)
        fresh = sums / counts.masked_fill(empty, 1)[:, None]
        means = torch.where(empty[:, None], means, fresh)
    return means, counts


class EuclideanCodebook(nn.Module):
    def __init__(self, dim: int, codebook_size: int, kmeans_init: int = False, kmeans_iters: int = 10,
                 decay: float = 0.99, epsilon: float = 1e-5, threshold_ema_dead_code: int = 2):
        super().__init__()
        self.decay = decay
        self.codebook_size = codebook_size
        self.kmeans_iters = kmeans_iters
        self.epsilon = epsilon
        self.threshold_ema_dead_code = threshold_ema_dead_code
        if kmeans_init:
            embed = torch.zeros(codebook_size, dim)
        else:
            embed = torch.empty(codebook_size, dim)
            nn.init.kaiming_uniform_(embed)
        self.register_buffer("inited", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(codebook_size))
        self.register_buffer("embed", embed)
        self.register_buffer("embed_avg", embed.clone())

    # -- state updates (training only) ---------------------------------------------------------
    def _init_from(self, rows):
        centres, counts = kmeans(rows, self.codebook_size, self.kmeans_iters)
        self.embed.data.copy_(centres)
        self.embed_avg.data.copy_(centres)
        self.cluster_size.data.copy_(counts)
        self.inited.data.fill_(1.0)

    def _revive_dead_codes(self, rows):
        if self.threshold_ema_dead_code == 0:
            return
        dead = self.cluster_size < self.threshold_ema_dead_code
        if not torch.any(dead):
            return
        self.embed.data.copy_(torch.where(dead[:, None], _pick_rows(rows, self.codebook_size), self.embed))

    def _ema_update(self, rows, onehot):
        d = self.decay
        self.cluster_size.data.mul_(d).add_(onehot.sum(0), alpha=1 - d)
        self.embed_avg.data.mul_(d).add_((rows.t() @ onehot).t(), alpha=1 - d)
        total = self.cluster_size.sum()
        smoothed = (self.cluster_size + self.epsilon) / (total + self.codebook_size * self.epsilon) * total
        self.embed.data.copy_(self.embed_avg / smoothed.unsqueeze(1))

    # -- API -------------------------------------------------------------------------------------
    def encode(self, x):
        idx = _nearest_code(x.reshape(-1, x.shape[-1]), self.embed)
        return idx.reshape(x.shape[0], x.shape[1])

    def decode(self, embed_ind):
        return F.embedding(embed_ind, self.embed)

    def forward(self, x):
        rows = x.reshape(-1, x.shape[-1])
        if not self.inited:
            self._init_from(rows)
        flat_idx = _nearest_code(rows, self.embed)
        idx = flat_idx.reshape(x.shape[0], x.shape[1])
        quantized = self.decode(idx)
        if self.training:
            self._revive_dead_codes(rows)
            self._ema_update(rows, F.one_hot(flat_idx, self.codebook_size).type(x.dtype))
        return quantized, idx


class VectorQuantization(nn.Module):
    def __init__(self, dim: int, codebook_size: int, codebook_dim: Optional[int] = None, decay: float = 0.99,
                 epsilon: float = 1e-5, kmeans_init: bool = True, kmeans_iters: int = 50,
                 threshold_ema_dead_code: int = 2, commitment_weight: float = 1.):
        super().__init__()
        cdim = codebook_dim or dim
        project = cdim != dim
        self.project_in = nn.Linear(dim, cdim) if project else nn.Identity()
        self.project_out = nn.Linear(cdim, dim) if project else nn.Identity()
        self.epsilon = epsilon
        self.commitment_weight = commitment_weight
        self._codebook = EuclideanCodebook(dim=cdim, codebook_size=codebook_size, kmeans_init=kmeans_init,
                                           kmeans_iters=kmeans_iters, decay=decay, epsilon=epsilon,
                                           threshold_ema_dead_code=threshold_ema_dead_code)
        self.codebook_size = codebook_size

    @property
    def codebook(self):
        return self._codebook.embed

    def encode(self, x):
        return self._codebook.encode(self.project_in(x.permute(0, 2, 1)))

    def decode(self, embed_ind):
        return self.project_out(self._codebook.decode(embed_ind)).permute(0, 2, 1)

    def forward(self, x):
        xt = self.project_in(x.permute(0, 2, 1))
        quantized, idx = self._codebook(xt)
        loss = torch.tensor([0.0], device=x.device, requires_grad=self.training)
        if self.training:
            quantized = xt + (quantized - xt).detach()            # straight-through estimator
            if self.commitment_weight > 0:
                loss = loss + F.mse_loss(quantized.detach(), xt) * self.commitment_weight
        return self.project_out(quantized).permute(0, 2, 1), idx, loss


class ResidualVectorQuantization(nn.Module):
    """Algorithm 1 of SoundStream (arXiv:2107.03312): each stage quantises the previous residual."""

    def __init__(self, num_quantizers, **kwargs):
        super().__init__()
        self.layers = nn.ModuleList([VectorQuantization(**kwargs) for _ in range(num_quantizers)])

    def forward(self, x):
        total, residual = 0.0, x
        losses, indices = [], []
        for vq in self.layers:
            q, idx, loss = vq(residual)
            residual = residual - q
            total = total + q
            indices.append(idx)
            losses.append(loss)
        return total, torch.stack(losses, 0).sum(), torch.stack(indices, 1)

    def encode(self, x):
        residual, out = x, []
        for vq in self.layers:
            idx = vq.encode(residual)
            residual = residual - vq.decode(idx)
            out.append(idx)
        return torch.stack(out, 1)

    def decode(self, q_indices):
        total = torch.tensor(0.0, device=q_indices.device)
        for i, vq in enumerate(self.layers):
            total = total + vq.decode(q_indices[:, i])
        return total
