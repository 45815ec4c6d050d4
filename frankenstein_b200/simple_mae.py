"""simple_mae -- host-side mirror of the reference ``models/simple_mae`` (time-bin tokens, RMSNorm blocks).

Same class names, constructor arguments, ``forward`` signatures and state-dict keys as the reference file
(models/simple_mae:1-407; the configs live in notebooks/simple_mae.ipynb cell 1).  Differences underneath:
RMSNorm is one fused kernel instead of five eager ops, attention is the label-mask flash kernel (the
``[B,1,T,T]`` padding masks of simple_mae:349-352 are never built), RoPE takes ``rope[:T]`` positions as in
simple_mae:39-42, and the per-forward debug ``print`` (simple_mae:379) is dropped.  A query row whose keys are
all padded yields zeros (the reference's math path yields NaN there); such rows never reach the loss.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .brainformer import (MLP, CausalSelfAttention, RMSNorm, _ConfigBase, _KernelLayerNorm, _linear_bf16,
                          build_complex_rope_cache)
from .ops import LabelMask, RopeSpec


@dataclass
class SimpleEncoderConfig(_ConfigBase):
    """notebooks/simple_mae.ipynb cell 1."""
    block_size: int = 768
    patch_size: int = 128
    n_layers: int = 6
    dim: int = 256
    hidden_dim: int = 1024
    head_dim: int = 32
    n_heads: int = 4
    n_kv_heads: int = 4
    rope_theta: int = 10000


@dataclass
class SimpleMAEConfig(_ConfigBase):
    """notebooks/simple_mae.ipynb cell 1."""
    n_layers: int = 2
    dim: int = 256
    hidden_dim: int = 1024
    head_dim: int = 32
    n_heads: int = 8
    n_kv_heads: int = 8
    rope_theta: int = 10000


class Block(nn.Module):
    """pre-RMSNorm block (models/simple_mae:194-205)."""

    def __init__(self, config):
        super().__init__()
        self.ln_1 = RMSNorm(config.dim)
        self.attn = CausalSelfAttention(config)
        self.ln_2 = RMSNorm(config.dim)
        self.mlp = MLP(config)

    def forward(self, x, attn_mask=None, rope=None, kv_cache=False):
        x = x + self.attn(self.ln_1(x, out_dtype=torch.bfloat16), attn_mask, rope, kv_cache=kv_cache)
        x = x + self.mlp(self.ln_2(x, out_dtype=torch.bfloat16))
        return x


def create_attention_mask_from_padding(x, pad_value=0):
    """models/simple_mae:226-233 as labels: a token is padded when all of its features equal pad_value."""
    return LabelMask.padding((x == pad_value).all(dim=2))


class SimpleEncoder(nn.Module):
    """models/simple_mae:240-297: one token per time bin."""

    def __init__(self, config, verbose: bool = False):
        super().__init__()
        self.config = config
        self.transformer = nn.ModuleDict(dict(
            emb=nn.Linear(config.patch_size, config.dim),
            h=nn.ModuleList([Block(config) for _ in range(config.n_layers)]),
            ln_f=_KernelLayerNorm(config.dim),
        ))
        self.precompute_rope_cash = build_complex_rope_cache(dim=config.head_dim, seq_len=config.block_size,
                                                             theta=config.rope_theta)
        self.attn_mask = None           # the reference keeps an all-True [T,T] tensor here (plain attribute)
        self._rope_table = None
        if verbose:
            print("Encoder: number of parameters: %.2fM" % (self.get_num_params() / 1e6,))

    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    @property
    def rope_cache(self) -> torch.Tensor:
        if self.precompute_rope_cash.device != self.device:
            self.precompute_rope_cash = self.precompute_rope_cash.to(device=self.device)
        return self.precompute_rope_cash

    def rope_table(self) -> torch.Tensor:
        if self._rope_table is None or self._rope_table.device != self.device:
            self._rope_table = torch.view_as_real(self.rope_cache).float().contiguous()
        return self._rope_table

    def get_num_params(self):
        return sum(p.numel() for p in self.parameters())

    def forward(self, x, attn_mask=None, rope_cache=None):
        """x [B, T, patch_size]; attn_mask: LabelMask | None; rope_cache: RopeSpec | None (None -> rope[:T])."""
        rope = rope_cache if rope_cache is not None else RopeSpec(self.rope_table(), None, 0)
        h = _linear_bf16(x, self.transformer.emb.weight, self.transformer.emb.bias).float()
        for block in self.transformer.h:
            h = block(h, attn_mask=attn_mask, rope=rope)
        return self.transformer.ln_f(h, out_dtype=torch.float32)


class SimpleMAE(nn.Module):
    """models/simple_mae:301-407."""

    def __init__(self, encoder_config, mae_config, verbose: bool = False):
        super().__init__()
        self.encoder_config = encoder_config
        self.encoder = SimpleEncoder(encoder_config, verbose)
        self.dim = mae_config.dim
        self.decoder = nn.ModuleDict(dict(
            emb=nn.Linear(encoder_config.dim, mae_config.dim),
            h=nn.ModuleList([Block(mae_config) for _ in range(mae_config.n_layers)]),
        ))
        self.mask_token = nn.Parameter(torch.randn(mae_config.dim))
        self.decoder_pos_emb = nn.Embedding(encoder_config.block_size, mae_config.dim)
        self.to_signals = nn.Linear(mae_config.dim, encoder_config.patch_size)
        if verbose:
            print("MAE: number of parameters: %.2fM" % (self.get_num_params() / 1e6))

    def get_num_params(self, non_embedding=True):
        return sum(p.numel() for p in self.parameters())

    def get_masking_indices(self, masking_ratio, x):
        b, n_tokens, _ = x.shape
        num_masked = int(masking_ratio * n_tokens)
        order = torch.rand(b, n_tokens, device=x.device).argsort(dim=-1)
        masked, unmasked = order[:, :num_masked], order[:, num_masked:]
        return torch.sort(masked, dim=1)[0], torch.sort(unmasked, dim=1)[0]

    def forward(self, x, targets=None, date_info=None, masking_ratio=0.75, return_preds=False):
        b, t, c = x.shape
        masked_indices, unmasked_indices = self.get_masking_indices(masking_ratio, x)
        rows = torch.arange(b, device=x.device)[:, None]
        is_padded = (x == 0).all(dim=2)

        # ---- encoder on the kept bins: padding labels and rope positions gathered per sample ----
        kept_mask = LabelMask.padding(is_padded[rows, unmasked_indices])
        rope = RopeSpec(self.encoder.rope_table(), unmasked_indices, 0)
        tokens = self.encoder(x[rows, unmasked_indices], attn_mask=kept_mask, rope_cache=rope)

        # ---- decoder on all bins with the full padding mask, no rope ----
        dec = torch.zeros(b, t, self.dim, device=x.device, dtype=torch.float32)
        dec[rows, unmasked_indices] = _linear_bf16(tokens, self.decoder.emb.weight, self.decoder.emb.bias).float()
        dec[rows, masked_indices] = self.mask_token.float()
        dec = dec + self.decoder_pos_emb(torch.cat([unmasked_indices, masked_indices], 1))
        full_mask = LabelMask.padding(is_padded)
        for block in self.decoder.h:
            dec = block(dec, full_mask)
        pred_tokens = _linear_bf16(dec, self.to_signals.weight, self.to_signals.bias).float()

        # ---- MSE on masked, non-padded bins (fused as a masked mean: no nonzero() host sync) ----
        pm, xm = pred_tokens[rows, masked_indices], x[rows, masked_indices]
        valid = (~is_padded[rows, masked_indices]).to(pm.dtype)[..., None]
        recon_loss = (((pm - xm) ** 2) * valid).sum() / (valid.sum() * c)
        if return_preds:
            binary_mask = torch.zeros_like(x)
            binary_mask[rows, masked_indices] = 1
            recon = torch.zeros_like(x)
            recon[rows, masked_indices] = pm.to(x.dtype)
            recon[rows, unmasked_indices] = x[rows, unmasked_indices]
            return recon_loss, recon, binary_mask
        return recon_loss, None
