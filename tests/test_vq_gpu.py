"""GPU parity of the VQ path against the CPU oracle (oracle/vector_quantize_ref.py), through the C ABI.

Bars (BASELINE.json north_star): code indices >= 99.9 % identical to the fp32 oracle with every
mismatch's relative distance gap < 1e-3; quantize / loss / gradients / EMA codebooks within rtol 2e-2
(the kernels keep fp32 masters, so the observed error is far smaller).
"""
import copy

import pytest
import torch

from tests.helpers import index_agreement, make_vq_problem, random_vq_problem

pytestmark = pytest.mark.gpu

RTOL = 2e-2


def _mods():
    from frankenstein_b200 import vector_quantize as fvq
    from oracle import vector_quantize_ref as ovq
    return fvq, ovq


@pytest.mark.parametrize("N,K,D", [(256, 128, 64), (300, 200, 64), (1000, 1000, 128), (4096, 512, 64),
                                   (2048, 4096, 256), (777, 130, 192)])
def test_search_accumulators_match_bf16_gemm(N, K, D):
    """Raw tcgen05 accumulators == fp32 GEMM of the bf16-rounded operands (layout / descriptor check)."""
    fvq, _ = _mods()
    from frankenstein_b200._lib import lib, ptr, stream, check
    X, C = random_vq_problem(N, K, D, seed=N + K)
    Xd, Cd = X.cuda(), C.cuda()
    xn, xb, _ = fvq.prepare_input(Xd, False)
    cb, c2pad = fvq.prepare_codebook(Cd, False)
    Kpad = c2pad.numel()
    S = lib().fk_vq_search_slots(N, K, 148)
    cand_val = torch.empty(N, S, fvq.CAND, device="cuda")
    cand_idx = torch.empty(N, S, fvq.CAND, device="cuda", dtype=torch.int32)
    dbg = torch.full((N, Kpad), float("nan"), device="cuda")
    check(lib().fk_vq_search_debug(ptr(xb), ptr(cb), ptr(c2pad), N, K, xb.shape[1], 0, ptr(cand_val), ptr(cand_idx), S,
                                   148, ptr(dbg), stream()), "fk_vq_search_debug")
    torch.cuda.synchronize()
    ref = xb.float() @ cb.float().t()
    got = dbg[:, :K]
    err = (got - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item() + 1e-4, f"accumulator mismatch {err}"
    # candidates: global best of the approximate scores must be among the slots
    score = (c2pad[:K][None, :] - 2.0 * got)
    best = score.argmin(dim=1)
    flat_idx = cand_idx.view(N, -1)
    assert (flat_idx == best[:, None].int()).any(dim=1).all()
    # keys are the scores with the low mantissa bits replaced by the tile tag (<= 2^-11 relative)
    v = cand_val.view(N, -1).masked_fill(flat_idx < 0, float("inf"))
    assert torch.allclose(v.min(dim=1).values, score.min(dim=1).values, rtol=2e-3, atol=1e-3)
    # every candidate's key is (close to) the score of the code it names
    valid = flat_idx >= 0
    named = torch.gather(score, 1, flat_idx.clamp_min(0).long())
    assert torch.allclose(v[valid], named[valid], rtol=2e-3, atol=1e-3)


@pytest.mark.parametrize("N,K,D,ctas", [(1000, 1000, 64, 148), (3000, 700, 64, 37), (5000, 384, 128, 148), (600, 2000, 64, 5),
                                        (16384, 8192, 256, 148)])
def test_search_marks_unowned_slots(N, K, D, ctas):
    """The kernel itself writes 'no candidate' into slots no CTA owns (no host-side clear): poison the buffers first."""
    fvq, _ = _mods()
    from frankenstein_b200._lib import lib, ptr, stream, check
    X, C = random_vq_problem(N, K, D, seed=N + ctas)
    _, xb, _ = fvq.prepare_input(X.cuda(), False)
    cb, c2pad = fvq.prepare_codebook(C.cuda(), False)
    S = lib().fk_vq_search_slots(N, K, ctas)
    cand_val = torch.full((N, S, fvq.CAND), -1e30, device="cuda")
    cand_idx = torch.full((N, S, fvq.CAND), 123456789, device="cuda", dtype=torch.int32)
    check(lib().fk_vq_search(ptr(xb), ptr(cb), ptr(c2pad), N, K, xb.shape[1], 0, ptr(cand_val), ptr(cand_idx), S, ctas,
                             stream()), "fk_vq_search")
    torch.cuda.synchronize()
    assert ((cand_idx == -1) | ((cand_idx >= 0) & (cand_idx < K))).all(), "stale or out-of-range candidate index"
    score = c2pad[:K][None, :] - 2.0 * (xb.float() @ cb.float().t())
    flat = cand_idx.view(N, -1)
    assert (flat == score.argmin(dim=1)[:, None].int()).any(dim=1).all()


@pytest.mark.parametrize("cosine", [False, True])
@pytest.mark.parametrize("N,K,D,kind", [(4096, 512, 64, "clustered"), (4096, 512, 64, "random"),
                                        (16384, 8192, 256, "clustered"), (8192, 8192, 256, "random"),
                                        (1000, 333, 64, "random"), (5, 7, 64, "random"),
                                        (4096, 16384, 256, "random"), (2048, 65536, 256, "clustered"),    # cfg-5 sweep ends
                                        (16384, 8192, 256, "random"),                                      # cfg 4 size, Gaussian
                                        (4096, 1024, 256, "random"), (4096, 2048, 256, "random"),          # remaining sweep K
                                        (4096, 4096, 256, "random"), (2048, 32768, 256, "random")])
def test_eval_forward_matches_oracle(N, K, D, kind, cosine):
    fvq, ovq = _mods()
    if kind == "clustered":
        X, C = make_vq_problem(N, K, D, seed=1, cosine=cosine, planted_ties=4 if K >= 16 else 0)
    else:
        X, C = random_vq_problem(N, K, D, seed=2)
        if cosine:
            C = torch.nn.functional.normalize(C, dim=-1)
    ref = ovq.VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine).eval()
    ref._codebook.embed.copy_(C[None])
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine).cuda().eval()
    mine._codebook.embed.copy_(C[None].cuda())
    mine._mark_dirty()
    mine._kmeans_initted_host = True
    q_ref, i_ref, l_ref = ref(X[None])
    q, i, l = mine(X[None].cuda())
    assert i.dtype == torch.int64 and tuple(i.shape) == (1, N) and tuple(l.shape) == (1,)
    Xs = torch.nn.functional.normalize(X, dim=-1) if cosine else X
    agree, worst = index_agreement(i, i_ref, Xs, C, cosine)
    assert agree >= 0.999, f"index agreement {agree}"
    assert worst < 1e-3, f"worst mismatch gap {worst}"
    same = (i.cpu() == i_ref).reshape(-1)
    assert torch.allclose(q.cpu().reshape(N, D)[same], q_ref.reshape(N, D)[same], rtol=RTOL, atol=1e-6)
    assert float(l) == 0.0


@pytest.mark.parametrize("cosine", [False, True])
@pytest.mark.parametrize("B,T,K,D", [(4, 128, 64, 64), (8, 128, 512, 64), (4, 256, 1024, 256)])
def test_train_step_matches_oracle(B, T, K, D, cosine):
    """loss, STE output, input gradient and the EMA-updated codebook state after 3 training steps."""
    fvq, ovq = _mods()
    X, C = make_vq_problem(B * T, K, D, seed=3, cosine=cosine, noise=0.5)
    ref = ovq.VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                                threshold_ema_dead_code=0).train()
    ref._codebook.embed.copy_(C[None]); ref._codebook.embed_avg.copy_(C[None]); ref._codebook.cluster_size.fill_(1.0)
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                              threshold_ema_dead_code=0).cuda().train()
    mine.load_state_dict(ref.state_dict())
    mine._kmeans_initted_host = True
    g = torch.Generator().manual_seed(5)
    for step in range(3):
        x = (X + 0.05 * step * torch.randn(X.shape, generator=g)).view(B, T, D)
        w = torch.randn(B, T, D, generator=g)
        xr = x.clone().requires_grad_(True)
        q_ref, i_ref, l_ref = ref(xr)
        ((q_ref * w).sum() + 3.0 * l_ref.sum()).backward()
        xm = x.cuda().requires_grad_(True)
        q, i, l = mine(xm)
        ((q * w.cuda()).sum() + 3.0 * l.sum()).backward()
        Xs = torch.nn.functional.normalize(x.view(-1, D), dim=-1) if cosine else x.view(-1, D)
        agree, worst = index_agreement(i, i_ref, Xs, ref._codebook.embed[0], cosine)
        assert agree == 1.0 or (agree >= 0.999 and worst < 1e-3), (step, agree, worst)
        same = (i.cpu() == i_ref).reshape(-1)
        if agree < 1.0:
            # a documented near-tie flipped: this step is compared on the agreeing rows (the loss, a mean over all rows,
            # moves by at most the flipped rows' share), and the oracle's state is re-based on ours so that the
            # remaining steps are compared exactly again instead of being skipped
            n_flip = int((~same).sum())
            assert torch.allclose(l.cpu(), l_ref, rtol=RTOL + 2.0 * n_flip / same.numel(), atol=1e-7), (step, l, l_ref)
            assert torch.allclose(q.detach().cpu().reshape(-1, D)[same], q_ref.detach().reshape(-1, D)[same], rtol=RTOL, atol=1e-6)
            assert torch.allclose(xm.grad.cpu().reshape(-1, D)[same], xr.grad.reshape(-1, D)[same], rtol=RTOL, atol=1e-5)
            ref.load_state_dict({k: v.cpu() for k, v in mine.state_dict().items()})
            continue
        assert torch.allclose(l.cpu(), l_ref, rtol=RTOL, atol=1e-7), (step, l, l_ref)
        assert torch.allclose(q.detach().cpu(), q_ref.detach(), rtol=RTOL, atol=1e-6)
        assert torch.allclose(xm.grad.cpu(), xr.grad, rtol=RTOL, atol=1e-5), (xm.grad.cpu() - xr.grad).abs().max()
        for name in ("cluster_size", "embed_avg", "embed"):
            a, b = getattr(mine._codebook, name).cpu(), getattr(ref._codebook, name)
            assert torch.allclose(a, b, rtol=RTOL, atol=1e-5), (step, name, (a - b).abs().max())


@pytest.mark.parametrize("cosine", [False, True])
def test_ema_trajectory_ten_steps(cosine):
    """Codebook state after 1 and after 10 EMA steps (SURVEY 8c fixture list): clustered inputs with a clear margin, so no
    near-tie can flip an assignment and the two trajectories must stay together to fp32 round-off."""
    fvq, ovq = _mods()
    B, T, K, D = 4, 128, 128, 64
    X, C = make_vq_problem(B * T, K, D, seed=11, cosine=cosine, noise=0.1)
    ref = ovq.VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                                threshold_ema_dead_code=0).train()
    ref._codebook.embed.copy_(C[None]); ref._codebook.embed_avg.copy_(C[None]); ref._codebook.cluster_size.fill_(1.0)
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                              threshold_ema_dead_code=0).cuda().train()
    mine.load_state_dict(ref.state_dict())
    mine._kmeans_initted_host = True
    g = torch.Generator().manual_seed(13)
    for step in range(10):
        x = (X + 0.02 * torch.randn(X.shape, generator=g)).view(B, T, D)
        _, i_ref, l_ref = ref(x)
        _, i, l = mine(x.cuda())
        assert torch.equal(i.cpu(), i_ref), f"assignment differs at step {step}"
        if step in (0, 9):
            assert torch.allclose(l.cpu(), l_ref, rtol=1e-4, atol=1e-7)
            for name in ("cluster_size", "embed_avg", "embed"):
                a, b = getattr(mine._codebook, name).cpu(), getattr(ref._codebook, name)
                assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (step, name, (a - b).abs().max())


@pytest.mark.parametrize("cosine", [False, True])
def test_dead_code_reset_matches_oracle(cosine):
    """K >> distinct inputs: most codes expire; same injected replacement rows -> same codebook."""
    fvq, ovq = _mods()
    B, T, K, D = 2, 64, 256, 64
    g = torch.Generator().manual_seed(7)
    protos = torch.randn(8, D, generator=g)
    X = (protos[torch.randint(0, 8, (B * T,), generator=g)] + 0.01 * torch.randn(B * T, D, generator=g)).view(B, T, D)
    C = torch.randn(K, D, generator=g)
    if cosine:
        C = torch.nn.functional.normalize(C, dim=-1)
    rows = torch.randperm(B * T, generator=g)
    ref = ovq.VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                                threshold_ema_dead_code=2, sample_fn=lambda s, n, generator=None: rows[:n] if n <= rows.numel() else rows[torch.arange(n) % rows.numel()]).train()
    ref._codebook.embed.copy_(C[None]); ref._codebook.embed_avg.copy_(C[None]); ref._codebook.cluster_size.fill_(1.0)
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                              threshold_ema_dead_code=2).cuda().train()
    mine.load_state_dict(ref.state_dict())
    mine._kmeans_initted_host = True
    mine.sample_rows_override = rows
    q_ref, i_ref, l_ref = ref(X)
    q, i, l = mine(X.cuda())
    assert (i.cpu() == i_ref).all()
    n_exp = int(mine.last_n_expired.item())
    assert n_exp == int(ref.last_expired.sum()) and n_exp > K // 2
    for name in ("cluster_size", "embed_avg", "embed"):
        a, b = getattr(mine._codebook, name).cpu(), getattr(ref._codebook, name)
        assert torch.allclose(a, b, rtol=RTOL, atol=1e-5), (name, (a - b).abs().max())
    # structural invariants of the reset (SURVEY 7 "hard parts")
    exp = ref.last_expired
    assert (mine._codebook.cluster_size[0].cpu()[exp] == 2).all()
    # second step uses the refreshed tensor-core operand
    q2_ref, i2_ref, _ = ref(X)
    q2, i2, _ = mine(X.cuda())
    Xs = torch.nn.functional.normalize(X.view(-1, D), dim=-1) if cosine else X.view(-1, D)
    agree, worst = index_agreement(i2, i2_ref, Xs, ref._codebook.embed[0], cosine)
    assert agree >= 0.999 or worst < 1e-3


@pytest.mark.parametrize("cosine", [False, True])
def test_kmeans_init_matches_oracle(cosine):
    fvq, ovq = _mods()
    N, K, D = 2048, 64, 64
    X, _ = make_vq_problem(N, K, D, seed=11, cosine=False, noise=0.4)
    g = torch.Generator().manual_seed(3)
    init = torch.randperm(N, generator=g)[:K]
    ref = ovq.VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine, kmeans_init=True,
                                threshold_ema_dead_code=0, sample_fn=lambda s, n, generator=None: init[:n]).eval()
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine, kmeans_init=True,
                              threshold_ema_dead_code=0).cuda().eval()
    mine.kmeans_init_override = init
    q_ref, i_ref, _ = ref(X[None])
    q, i, _ = mine(X[None].cuda())
    assert float(mine._codebook.initted) == 1.0
    a, b = mine._codebook.embed.cpu(), ref._codebook.embed
    assert torch.allclose(a, b, rtol=RTOL, atol=1e-4), (a - b).abs().max()
    assert torch.allclose(mine._codebook.cluster_size.cpu(), ref._codebook.cluster_size)
    assert (i.cpu() == i_ref).float().mean() >= 0.999


def test_cpu_tensor_is_rejected():
    fvq, _ = _mods()
    from frankenstein_b200._lib import FkError
    m = fvq.VectorQuantize(dim=64, codebook_size=16)
    with pytest.raises(FkError):
        m(torch.zeros(1, 4, 64))


def test_full_size_properties_cfg4():
    """BASELINE cfg 4 size (N=16384, K=8192, D=256): size-independent properties instead of the oracle:
    idempotence (quantising a codeword returns itself) and optimality against a chunked fp32 scan on the GPU."""
    fvq, _ = _mods()
    N, K, D = 16384, 8192, 256
    X, C = make_vq_problem(N, K, D, seed=21, noise=0.6)
    mine = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25).cuda().eval()
    mine._codebook.embed.copy_(C[None].cuda()); mine._mark_dirty(); mine._kmeans_initted_host = True
    q, i, _ = mine(X[None].cuda())
    Xd, Cd = X.cuda(), C.cuda()
    best = torch.empty(N, dtype=torch.int64, device="cuda")
    for s in range(0, N, 2048):
        best[s:s + 2048] = torch.cdist(Xd[s:s + 2048], Cd).argmin(dim=1)
    agree = (best == i.view(-1)).float().mean().item()
    assert agree >= 0.999, agree
    q2, i2, _ = mine(Cd[None])
    assert (i2.view(-1) == torch.arange(K, device="cuda")).float().mean() >= 0.999
    assert torch.equal(q2.view(K, D)[i2.view(-1) == torch.arange(K, device="cuda")],
                       Cd[i2.view(-1) == torch.arange(K, device="cuda")])


def test_ema_stats_deterministic_and_exact():
    """K5 as a segmented sum in ascending row order: bit-identical from run to run (the fp32-atomic version was not) and
    equal to an fp64 index_add; covers short segments (<= 32 rows, in-warp rank sort), long ones (a hot code) and
    empty codes."""
    fvq, _ = _mods()
    N, K, D = 16384, 512, 256
    g = torch.Generator().manual_seed(5)
    xn = torch.randn(N, D, generator=g).cuda()
    ind = torch.randint(0, K, (N,), generator=g)
    ind[:3000] = 7                       # one hot code (long segment)
    ind[ind == 11] = 12                  # one empty code
    ind = ind.cuda()
    first = fvq.ema_stats(xn, ind, K)
    for _ in range(3):
        assert torch.equal(fvq.ema_stats(xn, ind, K), first)
    ref = torch.zeros(K, D, dtype=torch.float64, device="cuda").index_add_(0, ind, xn.double())
    assert torch.allclose(first[:K * D].view(K, D).double(), ref, rtol=1e-5, atol=1e-4)
    assert torch.equal(first[K * D:], torch.bincount(ind, minlength=K).float())
    assert first[K * D + 11] == 0 and (first[11 * D:12 * D] == 0).all()
    # degenerate assignments: every row on one code (row-scan path, ragged N), and segments right at the path boundaries
    for N2, K2, make in ((5003, 64, lambda n: torch.full((n,), 5)),
                         (4000, 8, lambda n: torch.cat([torch.zeros(32), torch.ones(33), torch.full((1024,), 2), torch.full((1025,), 3),
                                                        torch.full((n - 2114,), 4)]))):
        x2 = torch.randn(N2, 64, generator=g).cuda()
        i2 = make(N2).long()[torch.randperm(N2, generator=g)].cuda()
        s2 = fvq.ema_stats(x2, i2, K2)
        assert torch.equal(fvq.ema_stats(x2, i2, K2), s2)
        ref2 = torch.zeros(K2, 64, dtype=torch.float64, device="cuda").index_add_(0, i2, x2.double())
        assert torch.allclose(s2[:K2 * 64].view(K2, 64).double(), ref2, rtol=1e-5, atol=1e-3)
        assert torch.equal(s2[K2 * 64:], torch.bincount(i2, minlength=K2).float())


@pytest.mark.parametrize("cosine", [False, True])
def test_side_stream_ema_equals_in_stream(cosine):
    """The EMA all-reduce + finalize on a side stream (overlap_ema) leaves exactly the state of the in-stream order, step
    after step (the next search waits on the event), and state_dict() waits for a pending update."""
    fvq, _ = _mods()
    B, T, K, D = 4, 128, 256, 64
    X, C = make_vq_problem(B * T, K, D, seed=3, cosine=cosine, noise=0.5)
    mods = []
    for mode in (False, "always"):
        m = fvq.VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine,
                               threshold_ema_dead_code=2).cuda().train()
        m._codebook.embed.copy_(C[None].cuda()); m._codebook.embed_avg.copy_(C[None].cuda()); m._codebook.cluster_size.fill_(1.0)
        m._mark_dirty(); m._kmeans_initted_host = True
        m.overlap_ema = mode
        m.sample_rows_override = torch.arange(B * T)
        mods.append(m)
    g = torch.Generator().manual_seed(1)
    for step in range(4):
        x = (X + 0.05 * torch.randn(X.shape, generator=g)).view(B, T, D).cuda()
        outs = [m(x) for m in mods]
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
        sd0, sd1 = mods[0].state_dict(), mods[1].state_dict()
        for k in sd0:
            assert torch.equal(sd0[k], sd1[k]), (step, k)


def test_module_pickles_and_residual_vq():
    """torch.save(model) works (no lambda hooks); ResidualVQ (imported by models/vq_brain.py:6) quantises residuals."""
    import io
    fvq, ovq = _mods()
    m = fvq.VectorQuantize(dim=64, codebook_size=32).cuda()
    torch.save(m, io.BytesIO())
    torch.manual_seed(0)
    rvq = fvq.ResidualVQ(dim=64, num_quantizers=3, codebook_size=64, kmeans_init=False, commitment_weight=0.25).cuda().eval()
    x = torch.randn(2, 50, 64).cuda()
    q, ind, loss = rvq(x)
    assert q.shape == x.shape and ind.shape == (2, 50, 3) and loss.shape == (1, 3)
    # oracle: the same greedy residual loop over the fp32 reference quantiser with the same codebooks
    resid, acc = x.cpu(), torch.zeros_like(x.cpu())
    for li, layer in enumerate(rvq.layers):
        ref = ovq.VectorQuantizeRef(dim=64, codebook_size=64, commitment_weight=0.25).eval()
        ref._codebook.embed.copy_(layer._codebook.embed.cpu())
        qr, ir, _ = ref(resid)
        assert (ir == ind[..., li].cpu()).float().mean() >= 0.999
        resid, acc = resid - qr, acc + qr
    assert torch.allclose(q.cpu(), acc, rtol=RTOL, atol=1e-4)
    # residual norm shrinks layer by layer
    assert (x - q).norm() < x.norm()
