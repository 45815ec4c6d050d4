"""Drop-in for ``vector_quantize_pytorch.VectorQuantize`` as the reference uses it.

Reference call sites: import ``models/vq_brain.py:6``, constructor ``models/vq_brain.py:184-193``
(``dim, codebook_size, commitment_weight=0.25, channel_last=True, kmeans_init=True,
threshold_ema_dead_code=2, use_cosine_sim=<flag>``), call ``models/vq_brain.py:209`` / ``:233``:
``quantize [B,N,D] fp32, indices [B,N] int64, loss [1] = vq(x)``.  State-dict keys are the upstream
ones (``_codebook.{initted,cluster_size,embed_avg,embed}``) so ``safetensors`` checkpoints written
by ``utils/train_utils.py:172`` load unchanged.

Everything numeric runs in the sm_100a kernels of ``libfk_b200.so``: tcgen05/TMEM search, exact
fp32 re-score + gather + straight-through + commitment loss, EMA statistics / finalize /
dead-code reset, backward.  There is no PyTorch or CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

import math
import random
from typing import Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from ._lib import DTYPE_CODE, FkError, check, counters, lib, on_tensor_device, ptr, require_cuda, require_device, stream, timed

BN = 128  # code tile of the search kernel (c2 padding granularity)
CAND = 4  # candidates per (row, slot) written by the search kernel


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


# ------------------------------------------------------------------------------------------------
# thin functional wrappers over the C ABI (tensors in, tensors out)
# ------------------------------------------------------------------------------------------------
@on_tensor_device
def prepare_input(e: torch.Tensor, use_cosine: bool):
    """e [N, D] (f32/bf16/f16, contiguous) -> (xn fp32 [N, D], x_bf16 [N, Dp], inv_norm [N] | None)."""
    require_cuda(e)
    require_device()
    N, D = e.shape
    Dp = _round_up(D, 64)
    if e.dtype not in DTYPE_CODE:
        raise FkError(f"unsupported input dtype {e.dtype}")
    alias = (not use_cosine) and e.dtype == torch.float32
    xn = e if alias else torch.empty(N, D, device=e.device, dtype=torch.float32)
    xb = torch.empty(N, Dp, device=e.device, dtype=torch.bfloat16)
    inv_norm = torch.empty(N, device=e.device, dtype=torch.float32) if use_cosine else None
    check(lib().fk_vq_prepare_input(ptr(e), DTYPE_CODE[e.dtype], N, D, Dp, int(use_cosine),
                                    0 if alias else ptr(xn), ptr(xb), ptr(inv_norm), stream()), "fk_vq_prepare_input")
    return xn, xb, inv_norm


@on_tensor_device
def prepare_codebook(embed: torch.Tensor, use_cosine: bool, cb: Optional[torch.Tensor] = None,
                     c2pad: Optional[torch.Tensor] = None):
    require_cuda(embed)
    require_device()
    K, D = embed.shape
    Dp, Kpad = _round_up(D, 64), _round_up(K, BN)
    if cb is None:
        cb = torch.empty(K, Dp, device=embed.device, dtype=torch.bfloat16)
    if c2pad is None:
        c2pad = torch.empty(Kpad, device=embed.device, dtype=torch.float32)
    check(lib().fk_vq_prepare_codebook(ptr(embed), K, D, Dp, Kpad, int(use_cosine), ptr(cb), ptr(c2pad), stream()),
          "fk_vq_prepare_codebook")
    return cb, c2pad


@on_tensor_device
def search(xb: torch.Tensor, cb: torch.Tensor, c2pad: torch.Tensor, K: int, use_cosine: bool, max_ctas: int = 0):
    """tcgen05 search -> (cand_val [N,S,4] fp32 keys, cand_idx [N,S,4] int32 code indices, -1 = none)."""
    require_cuda(xb, cb, c2pad)
    require_device()
    N, Dp = xb.shape
    if Dp > 256:
        raise FkError("codebook dim > 256 is not supported by the search kernel")
    if max_ctas <= 0:
        max_ctas = torch.cuda.get_device_properties(xb.device).multi_processor_count
    S = lib().fk_vq_search_slots(N, K, max_ctas)
    cand_val = torch.empty(N, S, CAND, device=xb.device, dtype=torch.float32)
    cand_idx = torch.empty(N, S, CAND, device=xb.device, dtype=torch.int32)
    with timed("vq_search", 2.0 * N * K * Dp):
        check(lib().fk_vq_search(ptr(xb), ptr(cb), ptr(c2pad), N, K, Dp, int(use_cosine), ptr(cand_val), ptr(cand_idx), S,
                                 max_ctas, stream()), "fk_vq_search")
    return cand_val, cand_idx


@on_tensor_device
def finish(xn, embed, cand_val, cand_idx, use_cosine: bool, training: bool, commitment_weight: float,
           want_quantize: bool = True):
    """exact re-score + gather (+ straight-through, commitment loss) -> (indices int64 [N], quantize, loss [1])."""
    require_cuda(xn, embed, cand_val, cand_idx)
    N, D = xn.shape
    K = embed.shape[0]
    S = cand_val.shape[1]
    indices = torch.empty(N, device=xn.device, dtype=torch.int64)
    quantize = torch.empty(N, D, device=xn.device, dtype=torch.float32) if (want_quantize or training) else None
    loss = torch.zeros(1, device=xn.device, dtype=torch.float32)
    partials = torch.empty(lib().fk_vq_finish_partials(N), device=xn.device, dtype=torch.float32) if training else None
    check(lib().fk_vq_finish(ptr(xn), ptr(embed), ptr(cand_val), ptr(cand_idx), N, K, D, S, int(use_cosine),
                             int(training), float(commitment_weight), ptr(indices), ptr(quantize), ptr(loss),
                             ptr(partials), counters(_lib.CTR_VQ_FINISH), stream()), "fk_vq_finish")
    return indices, quantize, loss


@on_tensor_device
def ema_stats(xn, indices, K: int, extra_rows: int = 0) -> torch.Tensor:
    """packed [K*D + K (+ extra_rows*D)] fp32 = embed_sum || bins (|| room for dead-code candidates)."""
    N, D = xn.shape
    stats = torch.empty(K * D + K + extra_rows * D, device=xn.device, dtype=torch.float32)
    ws = torch.empty(lib().fk_vq_ema_stats_ws(N, K), device=xn.device, dtype=torch.int32)
    check(lib().fk_vq_ema_stats(ptr(xn), ptr(indices), N, K, D, ptr(stats), ptr(ws), stream()), "fk_vq_ema_stats")
    return stats


def dead_code_layout(K: int, N: int, world: int):
    """(rows contributed per rank, total candidate rows R) of the dead-code tail of the packed EMA buffer:
    R = min(K, N * world) rounded down to a multiple of `world`, at least one row per rank."""
    per_rank = max(1, min(K, N * world) // world)
    return per_rank, per_rank * world


class _VQFunction(torch.autograd.Function):
    """(quantize_out, indices, loss) = f(x); backward = fk_vq_backward (STE + commitment term + normalize Jacobian)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, x, vq):
        quantize, indices, loss, xn, inv_norm = vq._forward_impl(x)
        ctx.vq = vq
        ctx.in_dtype = x.dtype
        ctx.in_shape = x.shape
        ctx.save_for_backward(xn, quantize, inv_norm if inv_norm is not None else torch.empty(0, device=x.device))
        indices = indices.view(x.shape[:-1])
        ctx.mark_non_differentiable(indices)
        return quantize.view(x.shape), indices, loss

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_q, _g_ind, g_loss):
        xn, quantize, inv_norm = ctx.saved_tensors
        vq = ctx.vq
        N, D = xn.shape
        if g_q is not None:
            g_q = g_q.reshape(N, D).contiguous().float()
        if g_loss is not None:
            g_loss = g_loss.reshape(1).contiguous().float()
        ge = torch.empty(N, D, device=xn.device, dtype=torch.float32)
        check(lib().fk_vq_backward(ptr(g_q), ptr(g_loss), ptr(xn), ptr(quantize),
                                   ptr(inv_norm) if vq.use_cosine_sim else 0, N, D, int(vq.use_cosine_sim),
                                   float(vq.commitment_weight), ptr(ge), stream()), "fk_vq_backward")
        return ge.view(ctx.in_shape).to(ctx.in_dtype), None


class _Codebook(nn.Module):
    """State container with the upstream buffer names."""

    def __init__(self, dim: int, codebook_size: int, use_cosine_sim: bool, kmeans_init: bool):
        super().__init__()
        if kmeans_init:
            embed = torch.zeros(1, codebook_size, dim)
        else:
            embed = torch.empty(1, codebook_size, dim)
            nn.init.kaiming_uniform_(embed)           # upstream `uniform_init`
            if use_cosine_sim:
                embed = torch.nn.functional.normalize(embed, p=2, dim=-1)
        self.register_buffer("initted", torch.Tensor([not kmeans_init]))
        self.register_buffer("cluster_size", torch.zeros(1, codebook_size))
        self.register_buffer("embed_avg", embed.clone())
        self.register_buffer("embed", embed)


class VectorQuantize(nn.Module):
    def __init__(self, dim, codebook_size, commitment_weight=1.0, channel_last=True, kmeans_init=False,
                 kmeans_iters=10, threshold_ema_dead_code=0, use_cosine_sim=False, decay=0.8, eps=1e-5,
                 sync_codebook=None, dead_code_sampling="affine", **unsupported):
        super().__init__()
        if unsupported:
            raise NotImplementedError(f"VectorQuantize options not used by the reference are not built: {sorted(unsupported)}")
        if not channel_last:
            raise NotImplementedError("the reference only uses channel_last=True")
        if dim % 4 != 0 or dim > 256:
            raise NotImplementedError("codebook dim must be a multiple of 4 and <= 256")
        self.dim = dim
        self.codebook_size = codebook_size
        self.commitment_weight = float(commitment_weight)
        self.kmeans_iters = kmeans_iters
        self.threshold_ema_dead_code = threshold_ema_dead_code
        self.use_cosine_sim = bool(use_cosine_sim)
        self.decay = float(decay)
        self.eps = float(eps)
        self.sync_codebook = sync_codebook        # None = decide per call from torch.distributed
        self.dead_code_sampling = dead_code_sampling
        self._codebook = _Codebook(dim, codebook_size, use_cosine_sim, kmeans_init)
        self._cb = None          # bf16 tensor-core operand [K, Dp]
        self._c2pad = None       # [Kpad]
        self._operand_dirty = True
        self._kmeans_initted_host = not kmeans_init
        self.last_n_expired = None
        self.last_bins = None
        # test hooks: inject the draws upstream takes from the global RNG
        self.sample_rows_override = None     # int64 [>=R] rows used for dead-code replacement
        self.kmeans_init_override = None     # int64 [K] rows used as initial k-means means
        # data-parallel EMA: the all-reduce of the packed statistics and the finalize kernels run on a side stream and
        # overlap the decoder forward and the whole backward (the updated codebook is first needed by the
        # NEXT step's search); the main stream only waits for them when the codebook is touched again.
        # True = when the statistics are all-reduced (world > 1), "always" = also on one GPU (tests), False = in stream.
        self.overlap_ema = True
        self._side_stream = None
        self._pending = None         # (event, tensors kept alive until the side stream has consumed them)
        self.register_load_state_dict_post_hook(VectorQuantize._after_load)
        self.register_state_dict_pre_hook(VectorQuantize._before_state_dict)

    # ---- operand cache ----------------------------------------------------------------------
    @staticmethod
    def _after_load(module, incompatible_keys):      # (a plain function, not a lambda: the module stays picklable)
        module._mark_dirty()

    @staticmethod
    def _before_state_dict(module, prefix, keep_vars):
        module._wait_pending()

    def _wait_pending(self):
        """Make the current stream wait for an EMA update still running on the side stream."""
        if self._pending is not None:
            torch.cuda.current_stream().wait_event(self._pending[0])
            self._pending = None

    def _mark_dirty(self):
        self._wait_pending()
        self._operand_dirty = True
        self._kmeans_initted_host = None      # re-read `initted` from the buffer

    def _apply(self, fn, *a, **k):
        self._mark_dirty()
        return super()._apply(fn, *a, **k)

    def _operands(self):
        self._wait_pending()
        embed = self._codebook.embed[0]
        if self._operand_dirty or self._cb is None or self._cb.device != embed.device:
            self._cb, self._c2pad = prepare_codebook(embed.contiguous(), self.use_cosine_sim)
            self._operand_dirty = False
        return self._cb, self._c2pad

    @property
    def codebook(self):
        self._wait_pending()
        return self._codebook.embed[0]

    def _sync(self) -> bool:
        if self.sync_codebook is not None:
            return bool(self.sync_codebook) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _is_initted(self) -> bool:
        if self._kmeans_initted_host is None:
            self._kmeans_initted_host = bool(self._codebook.initted.item())   # one sync after load/move only
        return self._kmeans_initted_host

    # ---- dead-code replacement rows ----------------------------------------------------------
    def _sample_rows(self, N: int, R: int, device) -> torch.Tensor:
        if self.sample_rows_override is not None:
            return self.sample_rows_override.to(device=device, dtype=torch.int64)[:R].contiguous()
        if self.dead_code_sampling == "randperm":
            if N >= R:
                return torch.randperm(N, device=device)[:R]
            return torch.randint(0, N, (R,), device=device)
        # affine walk (a + i*b) mod N with gcd(b, N) = 1: R distinct rows (when R <= N) from two host draws,
        # no device-side permutation.  Uses torch's CPU generator so torch.manual_seed controls it.
        a = int(torch.randint(0, N, (1,)).item())
        b = int(torch.randint(1, max(N, 2), (1,)).item())
        while math.gcd(b, N) != 1:
            b = b + 1 if b + 1 < N else 1
        i = torch.arange(R, device=device, dtype=torch.int64)
        return (a + i * b) % N

    # ---- k-means init (upstream init_embed_ / kmeans) -----------------------------------------
    @torch.no_grad()
    def _kmeans_init(self, xn: torch.Tensor, xb: torch.Tensor):
        cbk = self._codebook
        N, D = xn.shape
        K = self.codebook_size
        Dp, Kpad = _round_up(D, 64), _round_up(K, BN)
        if self.kmeans_init_override is not None:
            idx = self.kmeans_init_override.to(device=xn.device, dtype=torch.int64)
        elif N >= K:
            idx = torch.randperm(N, device=xn.device)[:K]
        else:
            idx = torch.randint(0, N, (K,), device=xn.device)
        means = xn.index_select(0, idx).contiguous()
        sync = self._sync()
        if sync:
            dist.broadcast(means, src=0)
        cb, c2pad = prepare_codebook(means, self.use_cosine_sim)
        stats = None
        for _ in range(self.kmeans_iters):
            cand_val, cand_idx = search(xb, cb, c2pad, K, self.use_cosine_sim)
            ind, _, _ = finish(xn, means, cand_val, cand_idx, self.use_cosine_sim, False, 0.0, want_quantize=False)
            stats = ema_stats(xn, ind, K)
            if sync:
                dist.all_reduce(stats)
            check(lib().fk_vq_kmeans_update(ptr(stats), ptr(means), K, D, Dp, Kpad, int(self.use_cosine_sim), ptr(cb),
                                            ptr(c2pad), stream()), "fk_vq_kmeans_update")
        bins = stats[K * D:K * D + K] if stats is not None else torch.zeros(K, device=xn.device)
        cbk.embed.data.copy_(means[None])
        cbk.embed_avg.data.copy_((means * bins[:, None])[None])
        cbk.cluster_size.data.copy_(bins[None])
        cbk.initted.data.fill_(1.0)
        self._kmeans_initted_host = True
        self._cb, self._c2pad = cb, c2pad
        self._operand_dirty = False

    # ---- forward ------------------------------------------------------------------------------
    @torch.no_grad()
    def _forward_impl(self, x: torch.Tensor):
        cbk = self._codebook
        K, D = self.codebook_size, self.dim
        flat = x.detach().reshape(-1, D).contiguous()
        training = self.training
        self._wait_pending()             # the previous step's EMA update (side stream) must have landed
        xn, xb, inv_norm = prepare_input(flat, self.use_cosine_sim)
        if not self._is_initted():
            self._kmeans_init(xn, xb)
        cb, c2pad = self._operands()
        cand_val, cand_idx = search(xb, cb, c2pad, K, self.use_cosine_sim)
        indices, quantize, loss = finish(xn, cbk.embed[0], cand_val, cand_idx, self.use_cosine_sim, training,
                                         self.commitment_weight if training else 0.0)
        if training:
            self._ema_step(xn, indices)     # after the gather: the search/gather use the pre-update codebook
        return quantize, indices, loss, xn, inv_norm

    @on_tensor_device
    def forward(self, x: torch.Tensor):
        require_cuda(x)
        require_device()
        if x.shape[-1] != self.dim:
            raise FkError(f"expected last dim {self.dim}, got {tuple(x.shape)}")
        if self.training and x.requires_grad and torch.is_grad_enabled():
            return _VQFunction.apply(x, self)
        quantize, indices, loss, _, _ = self._forward_impl(x)
        return quantize.view(x.shape), indices.view(x.shape[:-1]), loss

    @torch.no_grad()
    def _ema_step(self, xn: torch.Tensor, indices: torch.Tensor):
        cbk = self._codebook
        N, D = xn.shape
        K = self.codebook_size
        Dp, Kpad = _round_up(D, 64), _round_up(K, BN)
        sync = self._sync()
        world = dist.get_world_size() if sync else 1
        rank = dist.get_rank() if sync else 0
        thr = float(self.threshold_ema_dead_code)
        # dead-code candidates ride in the tail of the packed stats buffer so that ONE all-reduce carries
        # embed_sum, bins and the replacement rows (each rank fills its own slice, the SUM gathers them).
        R = 0
        if thr > 0:
            per_rank, R = dead_code_layout(K, N, world)
        stats = ema_stats(xn, indices, K, extra_rows=R)
        tail = None
        if R > 0:
            tail = stats[K * D + K:].view(R, D)
            if world > 1:
                tail.zero_()
            rows = self._sample_rows(N, per_rank, xn.device)
            torch.index_select(xn, 0, rows, out=tail[rank * per_rank:(rank + 1) * per_rank])
        ws_total = torch.empty(1, device=xn.device, dtype=torch.float32)
        ws_rank = torch.empty(K, device=xn.device, dtype=torch.int32)
        n_exp = torch.empty(1, device=xn.device, dtype=torch.int32)
        ar = torch.arange(R, device=xn.device, dtype=torch.int64) if R > 0 else None
        if self._cb is None or self._cb.device != xn.device:
            self._operands()

        def finalize():
            if sync:
                dist.all_reduce(stats)
            check(lib().fk_vq_ema_update(ptr(stats), ptr(cbk.cluster_size), ptr(cbk.embed_avg), ptr(cbk.embed), K, D, Dp, Kpad,
                                         int(self.use_cosine_sim), self.decay, self.eps, thr, ptr(ar), R, ptr(tail), R,
                                         ptr(self._cb), ptr(self._c2pad), ptr(ws_total), ptr(ws_rank), ptr(n_exp), stream()),
                  "fk_vq_ema_update")

        if self.overlap_ema == "always" or (self.overlap_ema and sync):
            if self._side_stream is None or self._side_stream.device != xn.device:
                self._side_stream = torch.cuda.Stream(device=xn.device)
            main = torch.cuda.current_stream()
            ready = torch.cuda.Event()
            ready.record(main)                       # statistics + candidate rows complete; gather/finish have read embed
            with torch.cuda.stream(self._side_stream):
                self._side_stream.wait_event(ready)
                finalize()
                done = torch.cuda.Event()
                done.record(self._side_stream)
            # the buffers were allocated on the main stream: keep them referenced until the main stream has waited on
            # `done` (so the caching allocator cannot hand them out while the side stream still uses them)
            self._pending = (done, (stats, tail, ws_total, ws_rank, n_exp, ar))
        else:
            finalize()
        self.last_bins = stats[K * D:K * D + K]
        self.last_n_expired = n_exp
        self._operand_dirty = False


class ResidualVQ(nn.Module):
    """``vector_quantize_pytorch.ResidualVQ`` (imported next to VectorQuantize at models/vq_brain.py:6, never wired by the
    reference): ``num_quantizers`` VectorQuantize layers applied greedily to the running residual.

    forward(x [B,N,D]) -> (quantized_out [B,N,D] = sum of the layers' outputs, indices [B,N,Q] int64, losses [1,Q]), as
    upstream: ``residual -= quantized.detach()`` between layers, so every layer's straight-through / commitment gradient
    reaches x and each layer keeps its own EMA codebook.  Every layer is the sm_100a search / finish / EMA path above.
    """

    def __init__(self, *, dim, num_quantizers, codebook_size, shared_codebook=False, **kwargs):
        super().__init__()
        self.num_quantizers = num_quantizers
        self.layers = nn.ModuleList([VectorQuantize(dim=dim, codebook_size=codebook_size, **kwargs)
                                     for _ in range(num_quantizers)])
        if shared_codebook:
            first = self.layers[0]._codebook
            for layer in self.layers[1:]:
                layer._codebook = first

    @property
    def codebooks(self):
        return torch.stack([layer.codebook for layer in self.layers], dim=0)

    def get_codes_from_indices(self, indices):
        """[B,N,Q] -> [Q,B,N,D] codewords."""
        return torch.stack([layer.codebook[indices[..., q]] for q, layer in enumerate(self.layers)], dim=0)

    def forward(self, x):
        quantized_out, residual = None, x
        all_indices, all_losses = [], []
        for layer in self.layers:
            quantized, indices, loss = layer(residual)
            residual = residual - quantized.detach()
            quantized_out = quantized if quantized_out is None else quantized_out + quantized
            all_indices.append(indices)
            all_losses.append(loss)
        return quantized_out, torch.stack(all_indices, dim=-1), torch.stack(all_losses, dim=-1)
