"""Encoder -> GPT-2 prefix hand-off (models/gpt2_model.py:178-196): oracle vs the reference fixture (CPU),
kernels vs oracle and fixture (GPU).  fp32 in / fp32 out is exact (one add per element); bf16 within 1e-2."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gpt2_prefix.pt")


def test_oracle_matches_reference_fixture():
    from oracle.gpt2_prefix_ref import embed_with_prefix
    g = torch.load(GOLD, weights_only=False)
    wte, wpe, prefix = (g[k].clone().requires_grad_(True) for k in ("wte", "wpe", "prefix"))
    x0 = embed_with_prefix(wte, wpe, g["idx"], prefix)
    assert torch.equal(x0.detach(), g["x0"])
    x0.backward(g["g"])
    assert torch.allclose(prefix.grad, g["dprefix"], rtol=0, atol=1e-7)
    assert torch.allclose(wpe.grad, g["dwpe"], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
def test_prefix_embed_matches_fixture_and_oracle():
    from frankenstein_b200.prefix import prefix_embed
    from oracle.gpt2_prefix_ref import embed_with_prefix
    g = torch.load(GOLD, weights_only=False)
    wte, wpe, prefix = (g[k].cuda().requires_grad_(True) for k in ("wte", "wpe", "prefix"))
    x0 = prefix_embed(wte, wpe, g["idx"].cuda(), prefix)
    assert torch.equal(x0.detach().cpu(), g["x0"])                       # bit exact in fp32
    x0.backward(g["g"].cuda())
    assert torch.equal(prefix.grad.cpu(), g["dprefix"])
    assert torch.allclose(wpe.grad.cpu(), g["dwpe"], rtol=1e-6, atol=1e-7)
    # wte is tied to lm_head in the reference (its recorded gradient mixes both uses): check the scatter against the oracle
    w2, p2, q2 = (g[k].clone().requires_grad_(True) for k in ("wte", "wpe", "prefix"))
    embed_with_prefix(w2, p2, g["idx"], q2).backward(g["g"])
    assert torch.allclose(wte.grad.cpu(), w2.grad, rtol=1e-5, atol=1e-6)
    # bf16 prefix in, bf16 out; no prefix; no tokens
    xb = prefix_embed(wte, wpe, g["idx"].cuda(), prefix.detach().to(torch.bfloat16), out_dtype=torch.bfloat16)
    assert torch.allclose(xb.float().cpu(), g["x0"], rtol=1e-2, atol=1e-2)
    x1 = prefix_embed(wte, wpe, g["idx"].cuda(), None)
    assert torch.equal(x1.detach().cpu(), embed_with_prefix(g["wte"], g["wpe"], g["idx"], None))


@pytest.mark.gpu
def test_prefix_embed_gpt2_size_and_errors():
    """GPT-2 small shapes: 128 trials x (32 prefix + 25 text tokens) x 768, vocabulary 50304."""
    from frankenstein_b200._lib import FkError
    from frankenstein_b200.prefix import embed_with_prefix as dev_embed
    from oracle.gpt2_prefix_ref import embed_with_prefix
    gen = torch.Generator().manual_seed(0)
    wte = torch.nn.Embedding(50304, 768)
    wpe = torch.nn.Embedding(1024, 768)
    idx = torch.randint(0, 50304, (128, 25), generator=gen)
    prefix = torch.randn(128, 32, 768, generator=gen)
    ref = embed_with_prefix(wte.weight.detach(), wpe.weight.detach(), idx, prefix)
    wte, wpe = wte.cuda(), wpe.cuda()
    out = dev_embed(wte, wpe, idx.cuda(), prefix.cuda())
    assert torch.equal(out.detach().cpu(), ref)
    with pytest.raises(FkError):
        dev_embed(wte, wpe, torch.full((2, 3), 50304, device="cuda"), None)          # id outside the vocabulary
    with pytest.raises(FkError):
        dev_embed(wte, wpe, idx.cuda()[:, :1].expand(128, 1000).contiguous(), prefix.cuda())   # longer than block_size
    with pytest.raises(FkError):
        dev_embed(torch.nn.Embedding(8, 4), torch.nn.Embedding(8, 4), torch.zeros(1, 2, dtype=torch.long))   # CPU tensors
