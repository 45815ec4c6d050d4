"""CPU fp32 restatement of ``vector_quantize_pytorch.VectorQuantize`` -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The arithmetic this file restates does not live in /root/reference: it is the
un-vendored PyPI dependency ``vector-quantize-pytorch`` (lucidrains), imported at
``models/vq_brain.py:6`` and constructed at ``models/vq_brain.py:184-193``.  The reference has
no requirements/lock file, so the version is unpinned (the repository's era, mid-2024, implies
~v1.14.x).  The package is not installed in this image and cannot be fetched (no network), and
the reference holds no test or golden vector for the quantiser (only the output *shape*
``[16,192,64]`` at ``notebooks_trainer/vq_brain_trainer.ipynb:46``).  This file therefore
restates the library's published algorithm (functions ``VectorQuantize.forward``,
``EuclideanCodebook.forward``, ``CosineSimCodebook.forward``, ``gumbel_sample`` deterministic
branch, ``cdist``, ``ema_inplace``, ``laplace_smoothing``, ``kmeans``, ``sample_vectors``,
``expire_codes_``/``replace``, ``l2norm``) and anchors parity on the reference's call sites:

* ctor kwargs            -- models/vq_brain.py:184-193
* ``quantizer(e)`` call   -- models/vq_brain.py:209 and :233
* returned triple         -- (quantize [B,N,D] fp32, indices [B,N] int64, loss [1])

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product path (``frankenstein_b200``) never does.

Everything the library draws from the global RNG (k-means seeds, dead-code replacement rows)
can be injected through ``sample_fn`` so that the CUDA path and this oracle can be driven with
identical draws.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import nn


def l2norm(t: torch.Tensor) -> torch.Tensor:
    # upstream: F.normalize(t, p=2, dim=-1)  (eps 1e-12)
    return F.normalize(t, p=2, dim=-1)


def cdist(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """upstream ``cdist``: sqrt(clamp(|x|^2 + |y|^2 - 2 x.y, 0)) on [h, n, d] x [h, c, d]."""
    x2 = (x ** 2).sum(-1)
    y2 = (y ** 2).sum(-1)
    xy = torch.einsum("hid,hjd->hij", x, y) * -2
    return (x2[..., :, None] + y2[..., None, :] + xy).clamp(min=0).sqrt()


def laplace_smoothing(x: torch.Tensor, n_categories: int, eps: float = 1e-5) -> torch.Tensor:
    denom = x.sum(dim=-1, keepdim=True)
    return (x + eps) / (denom + n_categories * eps)


def default_sample_fn(samples: torch.Tensor, num: int, generator=None) -> torch.Tensor:
    """upstream ``sample_vectors``: randperm when enough rows, randint (with repeats) otherwise.

    Returns the *indices* (the library returns ``samples[indices]``); returning indices lets a
    test feed the very same draw to the CUDA path.
    """
    n = samples.shape[0]
    if n >= num:
        return torch.randperm(n, generator=generator)[:num]
    return torch.randint(0, n, (num,), generator=generator)


def kmeans(samples: torch.Tensor, num_clusters: int, num_iters: int, use_cosine_sim: bool,
           init_indices: torch.Tensor, all_reduce_fn: Callable = lambda t: None):
    """upstream ``kmeans`` for one codebook head.  samples [n, d] -> (means [c, d], bins [c])."""
    means = samples[init_indices]
    bins = None
    for _ in range(num_iters):
        if use_cosine_sim:
            dists = samples @ means.t()
        else:
            dists = -cdist(samples[None], means[None])[0]
        buckets = torch.argmax(dists, dim=-1)
        bins = torch.bincount(buckets, minlength=num_clusters)
        all_reduce_fn(bins)
        zero_mask = bins == 0
        bins_min_clamped = bins.masked_fill(zero_mask, 1)
        new_means = torch.zeros(num_clusters, samples.shape[-1], dtype=samples.dtype, device=samples.device)
        new_means.index_add_(0, buckets, samples)
        new_means = new_means / bins_min_clamped[:, None]
        all_reduce_fn(new_means)
        if use_cosine_sim:
            new_means = l2norm(new_means)
        means = torch.where(zero_mask[:, None], means, new_means)
    return means, bins


class _Codebook(nn.Module):
    """State container with the upstream buffer names (``_codebook.*`` in the state_dict)."""

    def __init__(self, dim: int, codebook_size: int, use_cosine_sim: bool, kmeans_init: bool):
        super().__init__()
        if kmeans_init:
            embed = torch.zeros(1, codebook_size, dim)
        else:
            embed = torch.empty(1, codebook_size, dim)
            nn.init.kaiming_uniform_(embed)
            if use_cosine_sim:
                embed = l2norm(embed)
        self.register_buffer("initted", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("embed", embed)


class VectorQuantizeRef(nn.Module):
    """fp32 oracle with the ctor signature the reference uses (models/vq_brain.py:184-193)."""

    def __init__(self, dim, codebook_size, commitment_weight=1.0, channel_last=True,
                 kmeans_init=False, kmeans_iters=10, threshold_ema_dead_code=0,
                 use_cosine_sim=False, decay=0.8, eps=1e-5, sync_codebook=None,
                 sample_fn: Optional[Callable] = None, all_reduce_fn: Optional[Callable] = None):
        super().__init__()
        assert channel_last, "the reference only uses channel_last=True"
        self.dim = dim
        self.codebook_size = codebook_size
        self.commitment_weight = commitment_weight
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.reset_cluster_size = threshold_ema_dead_code
        self.use_cosine_sim = use_cosine_sim
        self.decay = decay
        self.eps = eps
        self.sample_fn = sample_fn or default_sample_fn
        self.all_reduce_fn = all_reduce_fn or (lambda t: None)
        self._codebook = _Codebook(dim, codebook_size, use_cosine_sim, kmeans_init)
        self.last_expired = None          # bool [K] mask of the codes replaced by the last call
        self.last_replacement_rows = None  # indices into the flattened batch used for replacement

    # --- upstream ``init_embed_`` ------------------------------------------------------------
    def _init_embed(self, flat: torch.Tensor):
        cb = self._codebook
        if bool(cb.initted.item()):
            return
        idx = self.sample_fn(flat, self.codebook_size)
        means, bins = kmeans(flat, self.codebook_size, self.kmeans_iters, self.use_cosine_sim,
                             idx, self.all_reduce_fn)
        bins = bins.to(flat.dtype)
        cb.embed.data.copy_(means[None])
        cb.embed_avg.data.copy_((means * bins[:, None])[None])
        cb.cluster_size.data.copy_(bins[None])
        cb.initted.data.copy_(torch.Tensor([True]))

    # --- upstream ``expire_codes_`` + ``replace`` ----------------------------------------------
    def _expire_codes(self, flat: torch.Tensor):
        self.last_expired = None
        self.last_replacement_rows = None
        if self.threshold_ema_dead_code == 0:
            return
        cb = self._codebook
        expired = cb.cluster_size[0] < self.threshold_ema_dead_code
        if not torch.any(expired):
            return
        samples = l2norm(flat) if self.use_cosine_sim else flat
        rows = self.sample_fn(samples, int(expired.sum().item()))
        sampled = samples[rows]
        cb.embed.data[0][expired] = sampled
        cb.cluster_size.data[0][expired] = self.reset_cluster_size
        cb.embed_avg.data[0][expired] = sampled * self.reset_cluster_size
        self.last_expired = expired.clone()
        self.last_replacement_rows = rows.clone()

    def search(self, flat: torch.Tensor) -> torch.Tensor:
        """distance + argmax (first maximum wins) -- the part the tcgen05 kernel replaces."""
        embed = self._codebook.embed[0]
        if self.use_cosine_sim:
            dist = flat @ embed.t()
        else:
            dist = -cdist(flat[None], embed[None])[0]
        return dist.argmax(dim=-1), dist

    def forward(self, x: torch.Tensor):
        cb = self._codebook
        shape = x.shape
        x = x.float()
        if self.use_cosine_sim:            # VectorQuantize.forward: x = codebook.transform_input(x)
            x = l2norm(x)
        flat = x.reshape(-1, shape[-1])
        with torch.no_grad():
            self._init_embed(flat.detach())
            ind, _ = self.search(flat.detach())
        embed = cb.embed[0]
        quantize = embed[ind].clone()      # training: onehot @ embed  == gather (same values)

        if self.training:
            with torch.no_grad():
                f = flat.detach()
                bins = torch.bincount(ind, minlength=self.codebook_size).to(f.dtype)
                self.all_reduce_fn(bins)
                cb.cluster_size.data[0].lerp_(bins, 1 - self.decay)
                embed_sum = torch.zeros(self.codebook_size, shape[-1], dtype=f.dtype, device=f.device)
                embed_sum.index_add_(0, ind, f)
                self.all_reduce_fn(embed_sum)
                cb.embed_avg.data[0].lerp_(embed_sum, 1 - self.decay)
                cs = cb.cluster_size[0]
                cluster_size = laplace_smoothing(cs, self.codebook_size, self.eps) * cs.sum(-1, keepdim=True)
                embed_normalized = cb.embed_avg[0] / cluster_size[:, None]
                if self.use_cosine_sim:
                    embed_normalized = l2norm(embed_normalized)
                cb.embed.data[0].copy_(embed_normalized)
                self._expire_codes(f)

        quantize = quantize.reshape(shape)
        ind = ind.reshape(shape[:-1])
        loss = torch.zeros(1, device=x.device, requires_grad=self.training)
        if self.training:
            commit_quantize = quantize.detach()
            quantize = x + (quantize - x).detach()
            if self.commitment_weight > 0:
                loss = loss + F.mse_loss(commit_quantize, x) * self.commitment_weight
        return quantize, ind, loss
