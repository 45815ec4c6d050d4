"""Few-query attention (perceiver resampler; models/brainformer.py:175-219, :247-268) against fp32 PyTorch
on the same bf16-rounded operands: forward within 2e-2 relative (bf16 output), gradients within 3e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(q, k, v, H, cache=None, q0=0, k0=0):
    from frankenstein_b200.brainformer import apply_rope
    B, Tq, W = q.shape
    S = k.shape[1]
    hd = W // H
    q4, k4, v4 = q.view(B, Tq, H, hd), k.view(B, S, H, hd), v.view(B, S, H, hd)
    if cache is not None:
        q4 = apply_rope(q4, cache[q0:q0 + Tq]).to(torch.bfloat16).float()
        k4 = apply_rope(k4, cache[k0:k0 + S]).to(torch.bfloat16).float()
    o = F.scaled_dot_product_attention(q4.transpose(1, 2), k4.transpose(1, 2), v4.transpose(1, 2))
    return o.transpose(1, 2).reshape(B, Tq, W)


def _close(a, b, tol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol}"


@pytest.mark.parametrize("B,H,hd,Tq,S", [(2, 4, 16, 32, 4096), (3, 2, 16, 25, 300), (1, 4, 32, 64, 1000), (2, 1, 64, 1, 257),
                                         (2, 4, 16, 32, 32), (1, 2, 32, 7, 5)])
@pytest.mark.parametrize("with_rope", [False, True])
def test_small_attention_matches_torch(B, H, hd, Tq, S, with_rope):
    from frankenstein_b200 import ops
    from frankenstein_b200.brainformer import build_complex_rope_cache
    g = torch.Generator().manual_seed(B * 1000 + S)
    W = H * hd
    q = (torch.randn(B, Tq, W, generator=g) * 1.5).to(torch.bfloat16).cuda()
    k = torch.randn(B, S, W, generator=g).to(torch.bfloat16).cuda()
    v = torch.randn(B, S, W, generator=g).to(torch.bfloat16).cuda()
    w = torch.randn(B, Tq, W, generator=g).to(torch.bfloat16).cuda()
    cache, spec, q0, k0 = None, None, 0, 0
    if with_rope:
        P = max(S, Tq) + 3
        cache = build_complex_rope_cache(hd, P, 10000.0).cuda()
        spec = ops.RopeSpec(torch.view_as_real(cache).float().contiguous(), None, 0)
        q0, k0 = P - Tq, P - S                                  # the reference's rope[-T:] convention for both sides
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
    out = ops.small_attention(qa, ka, va, H, rope=spec, rope_q0=q0, rope_k0=k0)
    out.backward(w)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _ref(qr, kr, vr, H, cache, q0, k0)
    ref.backward(w.float())
    _close(out, ref, 2e-2, "out")
    _close(qa.grad, qr.grad, 3e-2, "dq")
    _close(ka.grad, kr.grad, 3e-2, "dk")
    _close(va.grad, vr.grad, 3e-2, "dv")
    # deterministic (fixed-order chunk merge)
    out2 = ops.small_attention(q, k, v, H, rope=spec, rope_q0=q0, rope_k0=k0)
    assert torch.equal(out, out2)


def test_small_attention_rejects_unsupported():
    from frankenstein_b200 import ops
    from frankenstein_b200._lib import FkError
    q = torch.zeros(1, 65, 64, dtype=torch.bfloat16, device="cuda")
    k = torch.zeros(1, 128, 64, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(FkError):
        ops.small_attention(q, k, k, 4)                       # 65 queries
    with pytest.raises(FkError):
        ops.small_attention(q[:, :8], k, k, 8)                # head_dim 8
    with pytest.raises(FkError):
        ops.small_attention(q[:, :8].cpu(), k.cpu(), k.cpu(), 4)


def test_perceiver_gradients_match_fp32_modules():
    """BrainFormer's perceiver (cross-attention on the few-query kernel, self-attention Block with RoPE at head_dim 16)
    against the same modules evaluated by fp32 PyTorch ops through the dense compatibility path."""
    from frankenstein_b200 import brainformer as bf
    torch.manual_seed(0)
    cfg = bf.Config(encoder=bf.MAEConfig(window_size=64, n_electrodes=8, patch_size=8, dim=64, n_layers=1, head_dim=32, hidden_dim=128,
                                         n_heads=2, n_kv_heads=2), n_output_tokens=25, output_dim=48, dim=64, n_layers=2, head_dim=16,
                    hidden_dim=128, n_heads=4, n_kv_heads=4)
    m = bf.BrainFormer(cfg).cuda()
    torch.nn.init.normal_(m.learnable_queries, std=0.5)
    ctx = torch.randn(3, 64, 64, device="cuda")
    h0 = m.learnable_queries.expand(3, -1, -1)

    def run(dense):
        for p in m.parameters():
            p.grad = None
        h = h0
        for cb in m.perceiver.h:
            if dense:
                # dense compatibility path: an all-True mask tensor sends both attentions to library SDPA
                ca = torch.ones(1, 1, 25, 64, dtype=torch.bool, device="cuda")
                sa = torch.ones(1, 1, 25, 25, dtype=torch.bool, device="cuda")
                h = cb(h, ctx, sa, ca, sa_rope=m.rope_cache)
            else:
                h = cb(h, ctx, None, None, sa_rope=m.rope_cache)
        out = m.perceiver.ln_f(h, out_dtype=torch.float32)
        out.square().sum().backward()
        return out.detach(), {n: p.grad.clone() for n, p in m.perceiver.named_parameters() if p.grad is not None}

    o1, g1 = run(False)
    o2, g2 = run(True)
    _close(o1, o2, 3e-2, "perceiver out")
    assert g1.keys() == g2.keys() and len(g1) > 10
    for n in g1:
        _close(g1[n], g2[n], 6e-2, f"perceiver grad {n}")


@pytest.mark.parametrize("B,H,hd,Tq,S", [(2, 4, 16, 32, 4096), (3, 2, 16, 25, 300), (1, 4, 32, 64, 1000)])
def test_small_attention_fused_kv_equals_separate(B, H, hd, Tq, S):
    """k | v as ONE projection buffer read in place (CausalCrossAttention's fused projection): bit-identical to the
    separate-operand call, forward and gradients (d kv = [dk | dv])."""
    from frankenstein_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + S)
    W = H * hd
    q = (torch.randn(B, Tq, W, generator=g) * 1.5).to(torch.bfloat16).cuda()
    kv = torch.randn(B, S, 2 * W, generator=g).to(torch.bfloat16).cuda()
    w = torch.randn(B, Tq, W, generator=g).to(torch.bfloat16).cuda()
    qa, kva = q.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    oa = ops.small_attention_kv(qa, kva, H)
    oa.backward(w)
    qb = q.clone().requires_grad_(True)
    kb, vb = kv[..., :W].contiguous().requires_grad_(True), kv[..., W:].contiguous().requires_grad_(True)
    ob = ops.small_attention(qb, kb, vb, H)
    ob.backward(w)
    assert torch.equal(oa, ob)
    assert torch.equal(qa.grad, qb.grad)
    assert torch.equal(kva.grad[..., :W], kb.grad) and torch.equal(kva.grad[..., W:], vb.grad)
