"""CPU fp32 restatement of the reference Brainformer / MAE / simple_mae forward passes -- TEST INFRASTRUCTURE ONLY.

Plain functions over a state_dict with the reference's parameter names, written against
``models/brainformer.py`` and ``models/simple_mae`` of the reference: dense bool masks, complex-number RoPE and
``F.scaled_dot_product_attention`` exactly as the reference calls them.  It exists because the GPU box has no
``/root/reference``; ``scripts/make_golden.py`` pins it against the unmodified reference modules in the build
container (fixtures under ``tests/golden``) and ``tests/test_oracle_cpu.py`` re-checks it against them.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs (the CPU arm and the PyTorch-eager-on-GPU
comparison) import this; the functions follow the device of their inputs.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rope_cache(dim, seq_len, theta=10000, device=None):
    """build_complex_rope_cache (models/brainformer.py:56-68).  (`device`: the oracle is also run on the GPU in eager
    PyTorch, as the full-size checker of the parity tests and as bench.py's gpu_eager_baseline.)"""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
    ang = torch.outer(torch.arange(seq_len), freqs).float()
    return torch.polar(torch.ones_like(ang), ang).to(device)


def apply_rope(x, rope, last=True):
    """apply_rope: brainformer.py:70-91 takes rope[-T:] (last=True); simple_mae:30-52 takes rope[:T]."""
    T = x.size(1)
    if rope.dim() == 2:
        rope = rope[-T:] if last else rope[:T]
    else:
        rope = rope[:, -T:] if last else rope[:, :T]
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    return torch.view_as_real(xc * rope.unsqueeze(-2)).flatten(3).type_as(x)


def block_causal_mask(block_size, tok_per_time, device=None):
    """build_advanced_causal_mask (models/brainformer.py:93-111), loop form as in the reference."""
    mask = torch.tril(torch.ones(block_size, block_size))
    for i in range(0, block_size, tok_per_time):
        mask[i:i + tok_per_time, i:i + tok_per_time] = 1
    return mask.bool().to(device)


def self_attention(sd, pre, x, n_heads, attn_mask, rope, rope_last=True):
    """CausalSelfAttention.forward (models/brainformer.py:147-173)."""
    B, T, _ = x.shape
    q, k, v = (F.linear(x, sd[f"{pre}.{n}.weight"]).view(B, T, n_heads, -1) for n in ("qw", "kw", "vw"))
    if rope is not None:
        q, k = apply_rope(q, rope, rope_last), apply_rope(k, rope, rope_last)
    if attn_mask is not None:
        attn_mask = attn_mask[..., -T:, -T:]
    q, k, v = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if attn_mask is not None:
        # SDPA's math path turns rows with no visible key into NaN; the kernels define those rows as zero output
        dead = ~attn_mask.expand(B, 1, T, T).any(dim=-1, keepdim=True) if attn_mask.dim() == 4 else None
        res = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask)
        if dead is not None:
            res = torch.where(dead, torch.zeros_like(res), res)
    else:
        res = F.scaled_dot_product_attention(q, k, v)
    res = res.transpose(1, 2).reshape(B, T, -1)
    return F.linear(res, sd[f"{pre}.project.weight"])


def mlp(sd, pre, x):
    """MLP.forward (models/brainformer.py:123-124)."""
    return F.linear(F.silu(F.linear(x, sd[f"{pre}.w1.weight"])) * F.linear(x, sd[f"{pre}.w3.weight"]), sd[f"{pre}.w2.weight"])


def _norm(sd, pre, x, rms):
    if rms:   # RMSNorm (models/simple_mae:181-192)
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6)).type_as(x) * sd[f"{pre}.weight"]
    return F.layer_norm(x, (x.shape[-1],), sd[f"{pre}.weight"], sd[f"{pre}.bias"], 1e-5)


def block(sd, pre, x, n_heads, attn_mask=None, rope=None, rms=False, rope_last=True):
    """Block.forward (models/brainformer.py:242-245 / simple_mae:201-205)."""
    x = x + self_attention(sd, f"{pre}.attn", _norm(sd, f"{pre}.ln_1", x, rms), n_heads, attn_mask, rope, rope_last)
    return x + mlp(sd, f"{pre}.mlp", _norm(sd, f"{pre}.ln_2", x, rms))


def to_patches(x, p):
    """'b (t p1) c -> b (t c) p1' (models/brainformer.py:282)."""
    b, T, c = x.shape
    return x.view(b, T // p, p, c).permute(0, 1, 3, 2).reshape(b, (T // p) * c, p)


def n_layers(sd, prefix):
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))


def encoder_forward(sd, x, cfg, pre=""):
    """Encoder.forward (models/brainformer.py:333-352)."""
    xp = to_patches(x, cfg["patch_size"])
    n_tok = xp.shape[1]
    h = F.linear(xp, sd[pre + "transformer.emb.weight"], sd[pre + "transformer.emb.bias"])
    n_pat = cfg["window_size"] // cfg["patch_size"]
    h = h + sd[pre + "space_embedding"].repeat(1, n_pat, 1)[:, -n_tok:]
    block_size = n_pat * cfg["n_electrodes"]
    mask = block_causal_mask(block_size, cfg["n_electrodes"], x.device)
    rope = rope_cache(cfg["head_dim"], block_size, cfg.get("rope_theta", 10000), x.device)
    for i in range(n_layers(sd, pre + "transformer.h.")):
        h = block(sd, f"{pre}transformer.h.{i}", h, cfg["n_heads"], mask, rope)
    return F.layer_norm(h, (h.shape[-1],), sd[pre + "transformer.ln_f.weight"], sd[pre + "transformer.ln_f.bias"], 1e-5)


def mae_forward(sd, x, cfg, masked_indices, unmasked_indices):
    """MAE.forward (models/brainformer.py:415-486) with the masking indices supplied by the caller."""
    xp = to_patches(x, cfg["patch_size"])
    b, n_tok, _ = xp.shape
    dev = x.device
    rows = torch.arange(b, device=dev)[:, None]
    n_pat = cfg["window_size"] // cfg["patch_size"]
    block_size = n_pat * cfg["n_electrodes"]
    space = sd["encoder.space_embedding"].repeat(1, n_pat, 1).expand(b, -1, -1)[rows, unmasked_indices]
    rope = rope_cache(cfg["head_dim"], block_size, cfg.get("rope_theta", 10000), dev).expand(b, -1, -1)[rows, unmasked_indices]
    full = block_causal_mask(block_size, cfg["n_electrodes"], dev)
    sub = full.expand(b, -1, -1)[torch.arange(b, device=dev)[:, None, None], unmasked_indices[..., None], unmasked_indices[:, None, :]][:, None]
    tok = F.linear(xp[rows, unmasked_indices], sd["encoder.transformer.emb.weight"], sd["encoder.transformer.emb.bias"]) + space
    for i in range(n_layers(sd, "encoder.transformer.h.")):
        tok = block(sd, f"encoder.transformer.h.{i}", tok, cfg["n_heads"], sub, rope)
    tok = F.layer_norm(tok, (tok.shape[-1],), sd["encoder.transformer.ln_f.weight"], sd["encoder.transformer.ln_f.bias"], 1e-5)
    dec = torch.zeros(b, n_tok, tok.shape[-1], device=dev)
    dec[rows, unmasked_indices] = tok
    dec[rows, masked_indices] = sd["mask_token"]
    dec = dec + F.embedding(torch.cat([unmasked_indices, masked_indices], 1), sd["decoder_pos_emb.weight"])
    for i in range(n_layers(sd, "decoder.h.")):
        dec = block(sd, f"decoder.h.{i}", dec, cfg["n_heads"])
    pred = F.linear(dec[rows, masked_indices], sd["to_signals.weight"], sd["to_signals.bias"])
    return F.mse_loss(pred, xp[rows, masked_indices]), pred


def simple_mae_forward(sd, x, enc_cfg, dec_cfg, masked_indices, unmasked_indices):
    """SimpleMAE.forward (models/simple_mae:338-407) with the masking indices supplied by the caller."""
    b, t, c = x.shape
    dev = x.device
    rows = torch.arange(b, device=dev)[:, None]
    is_padded = (x == 0).all(dim=2)
    attn_mask = ~is_padded.unsqueeze(1) & ~is_padded.unsqueeze(2)
    sub = attn_mask[torch.arange(b, device=dev)[:, None, None], unmasked_indices[..., None], unmasked_indices[:, None, :]][:, None]
    rope = rope_cache(enc_cfg["head_dim"], enc_cfg["block_size"], enc_cfg.get("rope_theta", 10000), dev).expand(b, -1, -1)[rows, unmasked_indices]
    tok = F.linear(x[rows, unmasked_indices], sd["encoder.transformer.emb.weight"], sd["encoder.transformer.emb.bias"])
    for i in range(n_layers(sd, "encoder.transformer.h.")):
        tok = block(sd, f"encoder.transformer.h.{i}", tok, enc_cfg["n_heads"], sub, rope, rms=True, rope_last=False)
    tok = F.layer_norm(tok, (tok.shape[-1],), sd["encoder.transformer.ln_f.weight"], sd["encoder.transformer.ln_f.bias"], 1e-5)
    dec_tok = F.linear(tok, sd["decoder.emb.weight"], sd["decoder.emb.bias"])
    dec = torch.zeros(b, t, dec_tok.shape[-1], device=dev, dtype=dec_tok.dtype)     # (the reference allocates x.dtype, simple_mae:371, which index_put rejects under autocast; fp32 runs are unaffected)
    dec[rows, unmasked_indices] = dec_tok
    dec[rows, masked_indices] = sd["mask_token"].to(dec.dtype)
    dec = dec + F.embedding(torch.cat([unmasked_indices, masked_indices], 1), sd["decoder_pos_emb.weight"])
    for i in range(n_layers(sd, "decoder.h.")):
        dec = block(sd, f"decoder.h.{i}", dec, dec_cfg["n_heads"], attn_mask[:, None], None, rms=True)
    pred = F.linear(dec, sd["to_signals.weight"], sd["to_signals.bias"])
    pm, xm = pred[rows, masked_indices], x[rows, masked_indices]
    valid = ~is_padded[rows, masked_indices]
    return F.mse_loss(pm, xm, reduction="none")[valid.nonzero(as_tuple=True)].mean(), pred


def cross_attention(sd, pre, x, context, n_heads):
    """CausalCrossAttention.forward (models/brainformer.py:200-219), no mask."""
    B, T, _ = x.shape
    S = context.shape[1]
    q = F.linear(x, sd[f"{pre}.qw.weight"]).view(B, T, n_heads, -1).transpose(1, 2)
    k = F.linear(context, sd[f"{pre}.kw.weight"]).view(B, S, n_heads, -1).transpose(1, 2)
    v = F.linear(context, sd[f"{pre}.vw.weight"]).view(B, S, n_heads, -1).transpose(1, 2)
    res = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, T, -1)
    return F.linear(res, sd[f"{pre}.project.weight"])


def brainformer_forward(sd, x, enc_cfg, cfg, targets=None):
    """BrainFormer.forward (models/brainformer.py:532-558): encoder -> perceiver CrossBlocks -> ln_f -> to_motion -> L1."""
    ctx = encoder_forward(sd, x, enc_cfg, pre="encoder.")
    b = x.shape[0]
    h = sd["learnable_queries"].expand(b, -1, -1)
    rope = rope_cache(cfg["head_dim"], cfg["n_output_tokens"], cfg.get("rope_theta", 10000), x.device)
    for i in range(n_layers(sd, "perceiver.h.")):
        p = f"perceiver.h.{i}"
        # CrossBlock.forward (models/brainformer.py:257-268)
        h = h + cross_attention(sd, f"{p}.cross_attn", _norm(sd, f"{p}.ln_1", h, False), ctx, cfg["n_heads"])
        h = h + mlp(sd, f"{p}.mlp", _norm(sd, f"{p}.ln_2", h, False))
        h = block(sd, f"{p}.sa_block", h, cfg["n_heads"], None, rope)
    pred = F.linear(_norm(sd, "perceiver.ln_f", h, False), sd["perceiver.to_motion.weight"], sd["perceiver.to_motion.bias"])
    if targets is None:
        return None, pred
    return F.l1_loss(pred, targets), pred
