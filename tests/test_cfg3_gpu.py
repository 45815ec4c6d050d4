"""BASELINE config 3 at full size on the GPU: brainformer.MAE (75 % token masking: 1024 kept of 4096 tokens per trial,
decoder over all 4096) and SimpleMAE on x[B, 512, 256] (time-bin tokens), against the oracle restatement run in fp32
eager PyTorch on the same GPU with the same recorded masking indices.  Tolerance: bf16 compute / fp32 accumulate,
rtol 2e-2 on the loss and predictions, looser on parameter gradients (sums over 10^5..10^6 bf16 products)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def close(a, b, rtol, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    scale = b.abs().max().item() + 1e-12
    err = (a - b).abs().max().item()
    assert err <= rtol * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g}"


def _sd_grad(model):
    return {k: v.detach().clone().float().requires_grad_(v.is_floating_point() and "attn_mask" not in k)
            for k, v in model.state_dict().items()}


def test_mae_full_size_matches_oracle_on_gpu():
    """models/brainformer.py:415-486 at the BASELINE shape (S = 4096 tokens / trial, 1024 kept, gathered labels + rope)."""
    from frankenstein_b200 import brainformer as bf
    from oracle import brainformer_ref as oref
    from tests.helpers import synth_trials
    dev = torch.device("cuda")
    torch.manual_seed(0)
    cfg = dict(window_size=512, n_electrodes=256, patch_size=32, dim=512, n_layers=4, head_dim=32, hidden_dim=2048,
               n_heads=16, n_kv_heads=16, n_dec_layers=4, decoder_dim=512)
    mae = bf.MAE(bf.MAEConfig(**cfg)).to(dev)
    B = 2
    x = synth_trials(B, T=512, C=256, seed=5).to(dev)
    masked, unmasked = mae.get_masking_indices(0.75, mae.encoder.to_patches(x))
    assert unmasked.shape == (B, 1024) and masked.shape == (B, 3072)
    mae.get_masking_indices = lambda r, xx: (masked, unmasked)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = mae(x)
    loss.backward()
    sd = _sd_grad(mae)
    loss_ref, pred_ref = oref.mae_forward(sd, x, cfg, masked, unmasked)
    loss_ref.backward()
    close(loss, loss_ref, 2e-2, "MAE loss")
    p = dict(mae.named_parameters())
    # spot checks: 64 rows of the big matrices, whole small tensors
    for n in ("encoder.transformer.emb.weight", "encoder.space_embedding", "mask_token", "to_signals.weight",
              "encoder.transformer.h.0.attn.qw.weight", "encoder.transformer.h.3.mlp.w2.weight",
              "decoder.h.0.attn.kw.weight", "decoder.h.3.mlp.w1.weight", "decoder.h.3.ln_2.weight"):
        g, r = p[n].grad, sd[n].grad
        if g.dim() == 2 and g.shape[0] > 64:
            rows = torch.linspace(0, g.shape[0] - 1, 64).long()
            g, r = g[rows], r[rows.to(r.device)]
        close(g, r, 8e-2, f"MAE grad {n}")
    # reconstruction path on the same indices
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        _, recon, binary = mae(x, return_preds=True)
    rows = torch.arange(B, device=dev)[:, None]
    recon_tok = mae.encoder.to_patches(recon)
    close(recon_tok[rows, masked], pred_ref.detach(), 3e-2, "MAE masked-token predictions")
    assert torch.equal(recon_tok[rows, unmasked], mae.encoder.to_patches(x)[rows, unmasked])
    assert float(binary.mean()) == 0.75


def test_simple_mae_full_size_matches_oracle_on_gpu():
    """models/simple_mae:338-407 at x[B, 512, 256] (one token per time bin, padded tails, 75 % masking)."""
    from frankenstein_b200 import simple_mae as sm
    from oracle import brainformer_ref as oref
    from tests.helpers import synth_trials
    dev = torch.device("cuda")
    torch.manual_seed(1)
    ec = dict(block_size=512, patch_size=256, n_layers=4, dim=512, hidden_dim=2048, head_dim=32, n_heads=16, n_kv_heads=16,
              rope_theta=10000)
    mc = dict(n_layers=2, dim=512, hidden_dim=2048, head_dim=32, n_heads=16, n_kv_heads=16, rope_theta=10000)
    m = sm.SimpleMAE(sm.SimpleEncoderConfig(**ec), sm.SimpleMAEConfig(**mc)).to(dev)
    B = 8
    x = synth_trials(B, T=512, C=256, seed=9).to(dev)          # zero-padded tails of 0..128 bins
    masked, unmasked = m.get_masking_indices(0.75, x)
    # the reference's SDPA gives NaN for a kept query whose keys are all padded; keep at least one real bin visible
    m.get_masking_indices = lambda r, xx: (masked, unmasked)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = m(x)
    loss.backward()
    sd = _sd_grad(m)
    loss_ref, pred_ref = oref.simple_mae_forward(sd, x, ec, mc, masked, unmasked)
    loss_ref.backward()
    assert torch.isfinite(loss_ref)
    close(loss, loss_ref, 2e-2, "SimpleMAE loss")
    p = dict(m.named_parameters())
    for n in ("encoder.transformer.emb.weight", "decoder.emb.weight", "mask_token", "to_signals.weight",
              "encoder.transformer.h.0.attn.vw.weight", "encoder.transformer.h.3.mlp.w3.weight", "decoder.h.1.attn.project.weight",
              "encoder.transformer.h.2.ln_1.weight"):
        g, r = p[n].grad, sd[n].grad
        if g.dim() == 2 and g.shape[0] > 64:
            rows = torch.linspace(0, g.shape[0] - 1, 64).long()
            g, r = g[rows], r[rows.to(r.device)]
        close(g, torch.nan_to_num(r), 8e-2, f"SimpleMAE grad {n}")
