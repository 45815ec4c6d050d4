"""The C ABI holds no device-side global state ("re-entrant per stream"): the work hand-out
counters of the persistent attention kernels and the last-block counters of the reductions are caller-owned words, one
set per (device, stream).  Also: the modules run under the reference trainer's fp16 autocast (utils/train_utils.py:96)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _attn_once(ops, qkv, w, mask, H):
    x = qkv.clone().requires_grad_(True)
    out = ops.attention_qkv(x * 1.0, H, None, mask)
    (out.float() * w).sum().backward()
    return out.detach(), x.grad.detach()


def test_attention_on_two_streams_concurrently():
    """Two forward+backward passes in flight on two streams (each stream has its own counter words): results are
    bit-identical to the same passes run one after the other."""
    from frankenstein_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    B, S, H = 8, 2048, 8
    probs = []
    for i in range(2):
        qkv = torch.randn(B, S, 3 * H * 32, generator=g).to(dev).to(torch.bfloat16)
        w = torch.randn(B, S, H * 32, generator=g).to(dev)
        probs.append((qkv, w, ops.LabelMask.block_causal(B, S, 128 << i, dev)))
    serial = [_attn_once(ops, q, w, m, H) for q, w, m in probs]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(3):
        res = [None, None]
        for i, st in enumerate(streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                res[i] = _attn_once(ops, *probs[i], H)
        torch.cuda.synchronize()
        for i in range(2):
            assert torch.equal(res[i][0], serial[i][0]), (rep, i, "forward differs under concurrency")
            assert torch.equal(res[i][1], serial[i][1]), (rep, i, "backward differs under concurrency")


def test_launch_after_a_rejected_launch():
    """A call that fails its argument checks leaves nothing behind: the next launch hands out every work item."""
    from frankenstein_b200 import _lib, ops
    from frankenstein_b200._lib import FkError, check, counters, lib, ptr, stream
    dev = torch.device("cuda")
    B, S, H = 2, 512, 4
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, S, 3 * H * 32, generator=g).to(dev).to(torch.bfloat16)
    w = torch.randn(B, S, H * 32, generator=g).to(dev)
    mask = ops.LabelMask.block_causal(B, S, 64, dev)
    good = _attn_once(ops, qkv, w, mask, H)
    out = torch.empty(B, S, H * 32, device=dev, dtype=torch.bfloat16)
    with pytest.raises(FkError):
        check(lib().fk_attn_forward_tc(ptr(qkv), ptr(qkv), ptr(qkv), ptr(out), 0, B, H, S, 64, 1, 1, 1, 1, 1, 1, 1, 1,
                                       0, 0, 0, 0, 0, 0, 1.0, counters(_lib.CTR_ATTN_FWD), stream()), "bad head_dim")
    with pytest.raises(FkError):
        check(lib().fk_attn_forward_tc(ptr(qkv), ptr(qkv), ptr(qkv), ptr(out), 0, B, H, S, 32, 1, 1, 1, 1, 1, 1, 1, 1,
                                       0, 0, 0, 0, 0, 0, 1.0, 0, stream()), "null counters")
    again = _attn_once(ops, qkv, w, mask, H)
    assert torch.equal(good[0], again[0]) and torch.equal(good[1], again[1])
    torch.cuda.synchronize()
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    assert int(_lib._counter_sets[key].abs().sum()) == 0, "counter words must be back to zero after every launch"


def test_train_step_under_fp16_autocast_with_grad_scaler():
    """utils/train_utils.py:35,96 runs mixed_precision=True = fp16 autocast (+ GradScaler in accelerate).  The modules
    compute in bf16 regardless of the autocast dtype; one scaled step must run, stay finite and agree with the bf16-autocast
    step on the same weights."""
    from frankenstein_b200 import brainformer as bf
    from frankenstein_b200.vq_brain import SoundStream
    dev = torch.device("cuda")
    torch.manual_seed(0)
    cfg = bf.MAEConfig(window_size=256, n_electrodes=16, patch_size=8, dim=64, n_layers=2, head_dim=32, hidden_dim=128,
                       n_heads=2, n_kv_heads=2, n_dec_layers=2, decoder_dim=64)
    per = bf.Config(encoder=cfg, n_output_tokens=8, output_dim=24, dim=64, n_layers=1, head_dim=16, hidden_dim=128, n_heads=4,
                    n_kv_heads=4)
    model = bf.BrainFormer(per).to(dev).train()
    vq = SoundStream(C=32, D=64, codebook_size=64, n_electrodes=32, use_cosine_sim=True).to(dev).train()
    x = torch.randn(2, 256, 16, device=dev)
    xv = torch.randn(2, 128, 32, device=dev)
    t = torch.randn(2, 8, 24, device=dev)
    losses = {}
    for dt in (torch.bfloat16, torch.float16):
        model.zero_grad(set_to_none=True)
        vq.zero_grad(set_to_none=True)
        scaler = torch.amp.GradScaler("cuda", enabled=dt == torch.float16)
        with torch.autocast("cuda", dtype=dt):
            l_bf, _ = model(x, t)
            l_vq, _ = vq(xv)
            mae_loss, _ = bf.MAE(cfg).to(dev)(x) if dt == torch.float16 else (torch.zeros((), device=dev), None)
            loss = l_bf + l_vq.sum() + 0.0 * mae_loss
        scaler.scale(loss).backward()
        grads = [p.grad for p in list(model.parameters()) + list(vq.parameters()) if p.grad is not None]
        assert grads and all(torch.isfinite(g_).all() for g_ in grads), dt
        losses[dt] = float(l_bf.detach())
    assert abs(losses[torch.float16] - losses[torch.bfloat16]) <= 2e-2 * abs(losses[torch.bfloat16]) + 1e-3, losses
