"""GPU: edge cases of the hot path -- ragged / tiny / maximum shapes, padding-only inputs, exact ties, invalid input."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _vq_eval(X, C, cosine=False):
    from frankenstein_b200 import vector_quantize as fvq
    K, D = C.shape
    m = fvq.VectorQuantize(dim=D, codebook_size=K, use_cosine_sim=cosine).cuda().eval()
    m._codebook.embed.copy_(C[None].cuda()); m._mark_dirty(); m._kmeans_initted_host = True
    q, i, l = m(X[None].cuda())
    return q[0].cpu(), i[0].cpu(), l.cpu()


@pytest.mark.parametrize("N,K,D", [(1, 1, 4), (1, 3, 64), (257, 129, 68), (3, 5000, 256), (513, 2, 8), (4097, 127, 192)])
def test_vq_ragged_and_tiny_shapes(N, K, D):
    """row / code / dim counts that are not multiples of the 256 x 128 x 64 tiles, down to a single row or code."""
    g = torch.Generator().manual_seed(N * 7 + K)
    X, C = torch.randn(N, D, generator=g), torch.randn(K, D, generator=g)
    q, i, _ = _vq_eval(X, C)
    ref = torch.cdist(X.double(), C.double()).argmin(dim=1)
    d_mine = (X - C[i]).norm(dim=1)
    d_ref = (X - C[ref]).norm(dim=1)
    assert (i == ref).float().mean() >= 0.999 or torch.allclose(d_mine, d_ref, rtol=1e-3)
    assert torch.equal(q, C[i])
    assert i.min() >= 0 and i.max() < K


def test_vq_exact_ties_lowest_index_wins():
    """duplicated codewords and rows equal to a codeword: torch.argmax semantics = first (lowest) index."""
    g = torch.Generator().manual_seed(0)
    K, D = 300, 64
    C = torch.randn(K, D, generator=g)
    C[200] = C[17]
    C[299] = C[17]
    C[150] = C[3]
    X = torch.stack([C[17], C[3], C[200] * 1.0, C[5]])
    for cosine in (False, True):
        Cn = F.normalize(C, dim=-1) if cosine else C
        _, i, _ = _vq_eval(X, Cn, cosine)
        assert i.tolist() == [17, 3, 17, 5], i.tolist()


def test_vq_zero_and_large_rows():
    g = torch.Generator().manual_seed(1)
    K, D = 64, 64
    C = torch.randn(K, D, generator=g)
    X = torch.zeros(5, D)
    X[1] = 1e4 * torch.randn(D, generator=g)
    X[2] = 1e-20
    q, i, _ = _vq_eval(X, C)
    ref = torch.cdist(X.double(), C.double()).argmin(dim=1)
    assert torch.isfinite(q).all()
    assert (i == ref).all() or torch.allclose((X - C[i]).norm(dim=1), (X - C[ref]).norm(dim=1), rtol=1e-3)


def test_vq_rejects_bad_input():
    from frankenstein_b200 import vector_quantize as fvq
    from frankenstein_b200._lib import FkError
    m = fvq.VectorQuantize(dim=64, codebook_size=16).cuda()
    with pytest.raises(FkError):
        m(torch.zeros(1, 4, 32, device="cuda"))            # wrong feature dim
    with pytest.raises(FkError):
        m(torch.zeros(1, 0, 64, device="cuda"))            # empty batch
    with pytest.raises(NotImplementedError):
        fvq.VectorQuantize(dim=512, codebook_size=16)       # dim > 256 is not built
    with pytest.raises(NotImplementedError):
        fvq.VectorQuantize(dim=64, codebook_size=16, heads=2)


def test_masked_l1_all_padding_is_nan_like_reference():
    from frankenstein_b200.vq_brain import SoundStream
    gt = torch.zeros(2, 8, 16)
    pred = torch.randn(2, 8, 16)
    l = SoundStream.custom_l1_loss(None, pred.cuda(), gt.cuda())
    assert torch.isnan(l)                                   # torch.mean over an empty selection
    gt[1, 3, 5] = 2.0                                       # exactly one valid row
    l = SoundStream.custom_l1_loss(None, pred.cuda(), gt.cuda())
    assert torch.allclose(l.cpu(), (pred[1, 3] - gt[1, 3]).abs().mean(), rtol=1e-5)


@pytest.mark.parametrize("impl", ["tc", "legacy"])
@pytest.mark.parametrize("B,S,H,kind", [(1, 1, 1, "none"), (2, 33, 2, "block"), (1, 1000, 3, "pad_all"), (1, 4100, 1, "block")])
def test_attention_edge_shapes(B, S, H, kind, impl, monkeypatch):
    """single token, sequences shorter than a tile, ragged tails, a fully padded sample (zero output rows)."""
    from frankenstein_b200 import ops
    monkeypatch.setattr(ops, "ATTN_BWD_IMPL", impl)
    monkeypatch.setattr(ops, "ATTN_FWD_IMPL", impl)
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B, S, 3 * H * 32, generator=g).cuda().to(torch.bfloat16)
    if kind == "none":
        mask = None
    elif kind == "block":
        mask = ops.LabelMask.block_causal(B, S, 7, qkv.device)
    else:
        pad = torch.ones(B, S, dtype=torch.bool)
        pad[:, : S // 3] = False if B > 1 else True         # B == 1: everything padded
        mask = ops.LabelMask.padding(pad.cuda())
    x = qkv.clone().requires_grad_(True)
    w = torch.randn(B, S, H * 32, generator=g).cuda()
    out = ops.attention_qkv(x * 1.0, H, None, mask)
    (out.float() * w).sum().backward()
    ref_in = qkv.float().clone().requires_grad_(True)
    q, k, v = ref_in.view(B, S, 3, H, 32).unbind(2)
    dense = mask.dense() if mask is not None else None
    r = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=dense)
    if dense is not None:
        r = torch.where(dense.any(-1, keepdim=True), r, torch.zeros_like(r))
    r = r.transpose(1, 2).reshape(B, S, H * 32)
    (r * w).sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(x.grad).all()
    assert (out.float() - r).abs().max() <= 3e-2 * (r.abs().max() + 1e-3)
    # gradients against fp32 SDPA autograd on the same bf16-rounded inputs (not only finite)
    gref = torch.nan_to_num(ref_in.grad)
    assert (x.grad.float() - gref).abs().max() <= 4e-2 * (gref.abs().max() + 1e-3)
    if kind == "pad_all":
        assert float(out.abs().max()) == 0.0 and float(x.grad.abs().max()) == 0.0


def test_norm_and_swiglu_tiny():
    from frankenstein_b200 import ops
    x = torch.randn(1, 4, device="cuda", requires_grad=True)
    w = torch.ones(4, device="cuda", requires_grad=True)
    b = torch.zeros(4, device="cuda", requires_grad=True)
    y = ops.layer_norm(x, w, b, 1e-5, torch.float32)
    y.sum().backward()
    assert torch.allclose(y, F.layer_norm(x.detach(), (4,), w.detach(), b.detach(), 1e-5), atol=1e-5)
    h = torch.randn(1, 16, device="cuda", dtype=torch.bfloat16)
    assert tuple(ops.swiglu(h).shape) == (1, 8)
