"""Generate tests/golden/*.pt from the UNMODIFIED reference modules (run in the build container only).

The reference (/root/reference) is imported through oracle/ref_shims.py (stubs for the four packages it
needs that are not installed; vector_quantize_pytorch -> oracle.vector_quantize_ref, parity unpinned).
Everything drawn from an RNG inside a forward pass (MAE masking) is recorded and stored with the fixture
so that the oracle and the CUDA path can be driven with the same draws.  Fixtures are small (< 1 MB each).
"""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle.ref_shims import load_reference

OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def record_indices(model):
    rec = {}
    orig = model.get_masking_indices

    def wrapped(ratio, x):
        m, u = orig(ratio, x)
        rec["masked"], rec["unmasked"] = m.clone(), u.clone()
        return m, u

    model.get_masking_indices = wrapped
    return rec


def grads_of(model, names):
    p = dict(model.named_parameters())
    return {n: p[n].grad.clone() for n in names}


def soundstream_fixture(vq_mod, cosine, C=32):
    """C = 32: the conv widths fall outside the tcgen05 GEMMs' granularity (cuDNN path of the mirror); C = 64: every
    convolution of the mirror runs as an implicit GEMM on the library's own kernels (conv.py)."""
    torch.manual_seed(10 + int(cosine) + (C - 32))
    D, K, E = 64, 64, 64
    m = quiet(vq_mod.SoundStream, C=C, D=D, codebook_size=K, n_electrodes=E, use_cosine_sim=cosine)
    cb = torch.randn(K, D) * 0.2
    if cosine:
        cb = torch.nn.functional.normalize(cb, dim=-1)
    q = m.quantizer._codebook
    q.embed.copy_(cb[None]); q.embed_avg.copy_(cb[None]); q.cluster_size.fill_(3.0); q.initted.fill_(1.0)
    m.quantizer.threshold_ema_dead_code = 0          # no RNG-dependent reset inside the fixture step
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn(3, 64, E)
    x[0, 50:] = 0
    x[2, 60:] = 0
    m.train()
    loss, o = m(x)
    loss.backward()
    with torch.no_grad():
        e = m.encoder(x)
    names = ["encoder.layers.0.weight", "encoder.layers.2.layers.0.layers.2.bias", "encoder.layers.6.weight",
             "decoder.layers.0.weight", "decoder.layers.4.layers.0.weight", "decoder.layers.6.bias"]
    return dict(config=dict(C=C, D=D, codebook_size=K, n_electrodes=E, use_cosine_sim=cosine), state_dict=sd0, x=x,
                loss=loss.detach(), o=o.detach(), enc_out=e.detach(), grads=grads_of(m, names),
                after={k: v.clone() for k, v in m.state_dict().items() if "quantizer" in k},
                perplexity_of_arange=m.calculate_perp(torch.arange(K)[None] % 7))


def brainformer_fixture(bf):
    torch.manual_seed(20)
    ec = dict(window_size=64, n_electrodes=16, patch_size=8, dim=64, n_layers=2, head_dim=32, hidden_dim=128, n_heads=2,
              n_kv_heads=2, n_dec_layers=2, decoder_dim=64)
    x = torch.randn(2, 64, 16)
    enc = quiet(bf.Encoder, bf.MAEConfig(**ec))
    y = enc(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    enc_names = ["transformer.emb.weight", "space_embedding", "transformer.h.0.attn.qw.weight", "transformer.h.1.mlp.w2.weight",
                 "transformer.h.0.ln_1.weight", "transformer.ln_f.bias"]
    out = dict(enc_config=ec, x=x, enc_state=enc.state_dict(), enc_out=y.detach(), enc_w=w, enc_grads=grads_of(enc, enc_names))
    mae = quiet(bf.MAE, bf.MAEConfig(**ec))
    rec = record_indices(mae)
    loss, _ = mae(x)
    loss.backward()
    out.update(mae_state={k: v.clone() for k, v in mae.state_dict().items()}, mae_loss=loss.detach(), mae_masked=rec["masked"],
               mae_unmasked=rec["unmasked"],
               mae_grads=grads_of(mae, ["mask_token", "to_signals.weight", "decoder.h.1.attn.vw.weight", "encoder.transformer.emb.weight"]))
    pc = dict(n_output_tokens=8, output_dim=24, dim=64, n_layers=2, head_dim=16, hidden_dim=64, n_heads=4, n_kv_heads=4)
    full = quiet(bf.BrainFormer, bf.Config(encoder=bf.MAEConfig(**ec), **pc))
    torch.nn.init.normal_(full.learnable_queries)
    t = torch.randn(2, 8, 24)
    l, p = full(x, t)
    out.update(per_config=pc, full_state={k: v.clone() for k, v in full.state_dict().items()}, targets=t, full_loss=l.detach(),
               full_pred=p.detach())
    # known answers
    out["mask_6_2"] = bf.build_advanced_causal_mask(6, 2)
    out["rope_8_5"] = torch.view_as_real(bf.build_complex_rope_cache(8, 5, 10000))
    xr = torch.arange(2 * 3 * 1 * 8, dtype=torch.float32).view(2, 3, 1, 8) / 10
    out["rope_in"], out["rope_out"] = xr, bf.apply_rope(xr, bf.build_complex_rope_cache(8, 5, 10000))
    big = quiet(bf.Encoder, bf.MAEConfig(window_size=768, patch_size=32))
    out["params_encoder_768_32"] = sum(p.numel() for p in big.parameters())
    return out


def simple_mae_fixture(sm):
    from dataclasses import dataclass

    @dataclass
    class EC:
        block_size: int = 64
        patch_size: int = 16
        n_layers: int = 2
        dim: int = 64
        hidden_dim: int = 128
        head_dim: int = 32
        n_heads: int = 2
        n_kv_heads: int = 2
        rope_theta: int = 10000

    @dataclass
    class MC:
        n_layers: int = 2
        dim: int = 64
        hidden_dim: int = 128
        head_dim: int = 32
        n_heads: int = 2
        n_kv_heads: int = 2
        rope_theta: int = 10000

    torch.manual_seed(30)
    m = quiet(sm.SimpleMAE, EC(), MC())
    x = torch.randn(2, 64, 16)
    x[1, 52:] = 0                      # zero-padded tail (is_padded rows)
    rec = record_indices(m)
    # the reference's math-path SDPA gives NaN for query rows with no visible key; keep every kept/decoded row
    # non-degenerate by making sure padded bins are masked out of the loss only (the fixture stores what it got)
    loss, _ = quiet(m, x, masking_ratio=0.5)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    # reconstruction with the same recorded indices (return_preds=True path, models/simple_mae:397-405)
    m2 = quiet(sm.SimpleMAE, EC(), MC())
    m2.load_state_dict(m.state_dict())
    m2.get_masking_indices = lambda r, xx: (rec["masked"], rec["unmasked"])
    with torch.no_grad():
        _, recon, binary = quiet(m2, x, masking_ratio=0.5, return_preds=True)
    return dict(enc_config=EC().__dict__, mae_config=MC().__dict__, state={k: v.clone() for k, v in m.state_dict().items()}, x=x,
                loss=loss.detach(), masked=rec["masked"], unmasked=rec["unmasked"], grads=grads, recon=recon, binary=binary)


def main():
    os.makedirs(OUT, exist_ok=True)
    vq_mod, bf, sm = load_reference()
    torch.save(soundstream_fixture(vq_mod, False), os.path.join(OUT, "soundstream_euclid.pt"))
    torch.save(soundstream_fixture(vq_mod, True), os.path.join(OUT, "soundstream_cosine.pt"))
    torch.save(soundstream_fixture(vq_mod, False, C=64), os.path.join(OUT, "soundstream_euclid_c64.pt"))
    torch.save(brainformer_fixture(bf), os.path.join(OUT, "brainformer_small.pt"))
    torch.save(simple_mae_fixture(sm), os.path.join(OUT, "simple_mae_small.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
