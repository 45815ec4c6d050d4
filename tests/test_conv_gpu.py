"""SoundStream convolutions as implicit GEMMs (models/vq_brain.py:22-159) against fp32 F.conv1d /
F.conv_transpose1d on the same bf16-rounded operands: outputs within 2e-2 of the output scale, gradients within 3e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    a, b = a.float(), b.float()
    err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
    assert err <= tol, f"{what}: relative error {err:.3e} > {tol}"


@pytest.mark.parametrize("B,T,Cin,Cout,k,s", [(3, 64, 64, 128, 3, 1), (2, 50, 128, 64, 5, 1), (2, 64, 64, 64, 4, 2), (1, 33, 64, 64, 4, 2),
                                             (4, 512, 256, 256, 3, 1), (2, 512, 512, 256, 5, 1), (2, 7, 64, 64, 3, 1)])
def test_causal_conv_matches_torch(B, T, Cin, Cout, k, s):
    from frankenstein_b200 import conv
    g = torch.Generator().manual_seed(T + Cin + k)
    x = torch.randn(B, T, Cin, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, generator=g) * (Cin * k) ** -0.5).cuda().requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).cuda().requires_grad_(True)
    xa = x.clone().requires_grad_(True)
    y = conv.causal_conv1d_cl(xa, w, b, s)
    T_out = (T - 1) // s + 1
    wout = torch.randn(B, T_out, Cout, generator=g).cuda().to(torch.bfloat16)
    y.backward(wout)
    xr = x.float().requires_grad_(True)
    wr, br = w.detach().to(torch.bfloat16).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ref = F.conv1d(F.pad(xr.transpose(1, 2), [k - 1, 0]), wr, br, stride=s).transpose(1, 2)
    ref.backward(wout.float())
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    _close(y, ref, 2e-2, "conv out")
    _close(xa.grad, xr.grad, 3e-2, "conv dx")
    _close(w.grad, wr.grad, 3e-2, "conv dw")
    _close(b.grad, br.grad, 3e-2, "conv db")


@pytest.mark.parametrize("B,T,Cin,Cout", [(3, 32, 64, 64), (2, 17, 128, 64), (4, 128, 256, 256)])
def test_causal_conv_transpose_matches_torch(B, T, Cin, Cout):
    from frankenstein_b200 import conv
    g = torch.Generator().manual_seed(T + Cin)
    x = torch.randn(B, T, Cin, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cin, Cout, 4, generator=g) * (Cin * 2) ** -0.5).cuda().requires_grad_(True)
    b = (torch.randn(Cout, generator=g) * 0.1).cuda().requires_grad_(True)
    xa = x.clone().requires_grad_(True)
    y = conv.causal_conv_transpose1d_cl(xa, w, b)
    wout = torch.randn(B, 2 * T, Cout, generator=g).cuda().to(torch.bfloat16)
    y.backward(wout)
    xr = x.float().requires_grad_(True)
    wr, br = w.detach().to(torch.bfloat16).float().requires_grad_(True), b.detach().clone().requires_grad_(True)
    ref = F.conv_transpose1d(xr.transpose(1, 2), wr, br, stride=2)[..., :-2].transpose(1, 2)       # trailing k - s samples trimmed
    ref.backward(wout.float())
    assert y.shape == ref.shape
    _close(y, ref, 2e-2, "convT out")
    _close(xa.grad, xr.grad, 3e-2, "convT dx")
    _close(w.grad, wr.grad, 3e-2, "convT dw")
    _close(b.grad, br.grad, 3e-2, "convT db")


def test_soundstream_stacks_own_convs_match_cudnn_path(monkeypatch):
    """The conv encoder and decoder (outputs and every parameter gradient) on the library's GEMMs against the same modules
    on cuDNN (which the fixture tests pin to the unmodified reference), and the whole VQ-VAE step's loss.  (The stacks are
    compared on their own: through the quantiser a bf16-level difference in the encoder output may flip a near-tie.)"""
    from frankenstein_b200 import vq_brain
    torch.manual_seed(3)
    m = vq_brain.SoundStream(C=64, D=64, codebook_size=128, n_electrodes=64, use_cosine_sim=False).cuda().train()
    assert m.encoder._own_ok() and m.decoder._own_ok()
    x = torch.randn(4, 128, 64, device="cuda")
    z = torch.randn(4, 32, 64, device="cuda")
    we, wd = torch.randn(4, 32, 64, device="cuda"), torch.randn(4, 128, 64, device="cuda")
    res = {}
    for impl in ("cudnn", "own"):
        monkeypatch.setattr(vq_brain, "CONV_IMPL", impl)
        for p in m.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            e = m.encoder(x)
            o = m.decoder(z)
        ((e.float() * we).sum() + (o.float() * wd).sum()).backward()
        grads = {n: p.grad.detach().float().clone() for n, p in m.named_parameters() if p.grad is not None}
        res[impl] = (e.detach().float().clone(), o.detach().float().clone(), grads)
    _close(res["own"][0], res["cudnn"][0], 3e-2, "encoder output")
    _close(res["own"][1], res["cudnn"][1], 3e-2, "decoder output")
    assert res["own"][2].keys() == res["cudnn"][2].keys() and len(res["own"][2]) > 40
    for n in res["own"][2]:
        _close(res["own"][2][n], res["cudnn"][2][n], 6e-2, f"grad {n}")
    # whole step: the losses agree, and the step trains
    q = m.quantizer
    cb = torch.randn(128, 64, device="cuda") * 0.3
    losses = {}
    for impl in ("cudnn", "own"):
        monkeypatch.setattr(vq_brain, "CONV_IMPL", impl)
        q._codebook.embed.copy_(cb[None]); q._codebook.embed_avg.copy_(cb[None]); q._codebook.cluster_size.fill_(3.0)
        q._codebook.initted.fill_(1.0); q._mark_dirty(); q._kmeans_initted_host = True
        q.threshold_ema_dead_code = 0
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, o = m(x)
        loss.sum().backward()
        assert o.shape == x.shape and torch.isfinite(o.float()).all()
        losses[impl] = float(loss.sum())
    assert abs(losses["own"] - losses["cudnn"]) <= 3e-2 * abs(losses["cudnn"]), losses


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,C,rpt,left,slack,view", [(3, 50, 64, 56, 4, 5, False), (1, 7, 8, 9, 2, 3, False), (4, 33, 128, 40, 0, 2, True),
                                                       (128, 512, 256, 514, 2, 3, False)])
def test_pad_rows_bit_exact(dtype, B, T, C, rpt, left, slack, view):
    """fk_pad_rows == zeros + strided copy (the F.pad of CausalConv1d, models/vq_brain.py:24-27), bit for bit."""
    from frankenstein_b200 import conv
    g = torch.Generator().manual_seed(B * 7 + T)
    if view:        # a row-strided view, as the sliced output of the previous convolution is
        base = torch.randn(B, T + 5, C, generator=g).cuda().to(dtype)
        x = base[:, :T]
    else:
        x = torch.randn(B, T, C, generator=g).cuda().to(dtype)
    got = conv._padded(x, rpt, left, slack)
    ref = torch.zeros(B * rpt + slack, C, device="cuda", dtype=torch.bfloat16)
    ref[:B * rpt].view(B, rpt, C)[:, left:left + T] = x
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    assert torch.equal(got, ref)


@pytest.mark.parametrize("M,N,ld", [(1000, 64, 64), (5, 256, 256), (70000, 512, 512), (40990, 320, 384), (5000, 4096, 4096), (65536, 1536, 1536)])
def test_column_sum(M, N, ld):
    """bias gradient = sum of dY over the tokens: fp32 sums of bf16 data, deterministic (two runs bit-identical)."""
    from frankenstein_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    full = torch.randn(M, ld, generator=g).cuda().to(torch.bfloat16)
    g2 = full[:, :N]
    a, b = ops.column_sum(g2), ops.column_sum(g2)
    assert torch.equal(a, b)
    ref = g2.double().sum(0)
    scale = g2.double().abs().sum(0).max().item()
    assert (a.double() - ref).abs().max().item() <= 1e-6 * scale
