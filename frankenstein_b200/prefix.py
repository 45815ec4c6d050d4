"""Encoder -> GPT-2 decoder hand-off: the first lines of the reference's ``GPT.forward``
(models/gpt2_model.py:178-196) --

    tok_emb = self.transformer.wte(idx)
    tok_emb = torch.cat([prefix, tok_emb], dim=1)
    x = self.transformer.drop(tok_emb + self.transformer.wpe(arange(t_ctx + t)))

as one fused kernel (forward) and one (backward), behind a function that takes the reference's own modules:
``embed_with_prefix(gpt.transformer.wte, gpt.transformer.wpe, idx, prefix)``.  The decoder itself is out of scope
(GPT-2 is the consumer of the path, not part of it).  No CPU path.
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import DTYPE_CODE, FkError, check, lib, on_tensor_device, ptr, require_cuda, require_device, stream


class _PrefixEmbedFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
    def forward(ctx, wte, wpe, idx, prefix, out_dtype):
        require_cuda(wte, wpe, idx, prefix)
        require_device()
        B, T = idx.shape
        Tc = 0 if prefix is None else prefix.shape[1]
        V, D = wte.shape
        if prefix is not None and (prefix.shape[0] != B or prefix.shape[2] != D):
            raise FkError(f"prefix {tuple(prefix.shape)} does not match batch {B} / n_embd {D}")
        if Tc + T > wpe.shape[0]:
            raise FkError(f"sequence of {Tc + T} tokens exceeds the position table ({wpe.shape[0]})")
        if T > 0 and (int(idx.min()) < 0 or int(idx.max()) >= V):
            raise FkError("token id outside the vocabulary")
        idx = idx.contiguous().to(torch.int64)
        pf = None
        if prefix is not None:
            pf = prefix.contiguous()
            if pf.dtype not in (torch.float32, torch.bfloat16):
                pf = pf.float()
        out = torch.empty(B, Tc + T, D, device=wte.device, dtype=out_dtype)
        wte_f, wpe_f = wte.detach().float().contiguous(), wpe.detach().float().contiguous()
        check(lib().fk_prefix_embed_forward(ptr(pf), 0 if pf is None else DTYPE_CODE[pf.dtype], ptr(idx), ptr(wte_f), ptr(wpe_f),
                                            ptr(out), DTYPE_CODE[out_dtype], B, Tc, T, D, V, wpe.shape[0], stream()),
              "fk_prefix_embed_forward")
        ctx.save_for_backward(idx)
        ctx.dims = (B, Tc, T, D, V, wpe.shape[0])
        ctx.dtypes = (wte.dtype, wpe.dtype, None if prefix is None else prefix.dtype)
        return out

    @staticmethod
    @on_tensor_device
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, Tc, T, D, V, P = ctx.dims
        g = g.contiguous().float()
        dwte = torch.zeros(V, D, device=g.device, dtype=torch.float32)
        dwpe = torch.zeros(P, D, device=g.device, dtype=torch.float32)
        check(lib().fk_prefix_embed_backward(ptr(g), ptr(idx), ptr(dwte), ptr(dwpe), B, Tc, T, D, V, stream()),
              "fk_prefix_embed_backward")
        dprefix = None
        if ctx.dtypes[2] is not None and ctx.needs_input_grad[3]:
            dprefix = g[:, :Tc].to(ctx.dtypes[2])
        return dwte.to(ctx.dtypes[0]), dwpe.to(ctx.dtypes[1]), None, dprefix, None


def prefix_embed(wte_weight: torch.Tensor, wpe_weight: torch.Tensor, idx: torch.Tensor, prefix: Optional[torch.Tensor] = None,
                 out_dtype=torch.float32) -> torch.Tensor:
    """[B, t_ctx + t, n_embd] = cat([prefix, wte[idx]], 1) + wpe[:t_ctx + t]  (models/gpt2_model.py:183-196)."""
    if not wte_weight.is_cuda:
        raise FkError("frankenstein_b200 kernels run on a B200 only (no CPU fallback)")
    return _PrefixEmbedFn.apply(wte_weight, wpe_weight, idx, prefix, out_dtype)


def embed_with_prefix(wte: torch.nn.Embedding, wpe: torch.nn.Embedding, idx, prefix=None, out_dtype=torch.float32):
    """Same through the reference's own modules: ``embed_with_prefix(gpt.transformer.wte, gpt.transformer.wpe, idx, prefix)``
    returns what ``GPT.forward`` feeds its first block (dropout p = 0, the reference's default)."""
    return prefix_embed(wte.weight, wpe.weight, idx, prefix, out_dtype)
