"""CPU: the oracle restatements (oracle/*.py) against the golden fixtures produced by the UNMODIFIED reference
(scripts/make_golden.py, build container).  These pin the oracle; the GPU tests then compare the CUDA path with
the oracle and with the same fixtures."""
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def sd_grad(sd):
    return {k: v.detach().clone().requires_grad_(v.is_floating_point() and "attn_mask" not in k) for k, v in sd.items()}


@pytest.mark.parametrize("name", ["soundstream_euclid.pt", "soundstream_cosine.pt"])
def test_soundstream_oracle_matches_reference(name):
    from oracle.soundstream_ref import SoundStreamRef, calculate_perp
    g = load(name)
    c = g["config"]
    m = SoundStreamRef(g["state_dict"], c["D"], c["codebook_size"], c["use_cosine_sim"], training=True)
    m.vq.threshold_ema_dead_code = 0
    loss, o = m(g["x"])
    loss.backward()
    assert tuple(loss.shape) == (1,)
    assert torch.allclose(loss.detach(), g["loss"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(o.detach(), g["o"], rtol=1e-4, atol=1e-5)
    for n, ref in g["grads"].items():
        assert torch.allclose(m.sd[n].grad, ref, rtol=1e-3, atol=1e-6), n
    for k, ref in g["after"].items():
        assert torch.allclose(m.vq.state_dict()[k[len("quantizer."):]], ref, rtol=1e-5, atol=1e-6), k
    K = c["codebook_size"]
    assert torch.allclose(calculate_perp(torch.arange(K)[None] % 7, K), g["perplexity_of_arange"])


def test_brainformer_oracle_matches_reference():
    from oracle import brainformer_ref as o
    g = load("brainformer_small.pt")
    ec = g["enc_config"]
    sd = sd_grad(g["enc_state"])
    y = o.encoder_forward(sd, g["x"], ec)
    (y * g["enc_w"]).sum().backward()
    assert torch.allclose(y.detach(), g["enc_out"], rtol=1e-4, atol=1e-5)
    for n, ref in g["enc_grads"].items():
        assert torch.allclose(sd[n].grad, ref, rtol=1e-3, atol=1e-5), n
    sd = sd_grad(g["mae_state"])
    loss, _ = o.mae_forward(sd, g["x"], ec, g["mae_masked"], g["mae_unmasked"])
    loss.backward()
    assert torch.allclose(loss.detach(), g["mae_loss"], rtol=1e-5)
    for n, ref in g["mae_grads"].items():
        assert torch.allclose(sd[n].grad, ref, rtol=1e-3, atol=1e-6), n
    l, p = o.brainformer_forward(g["full_state"], g["x"], ec, g["per_config"], g["targets"])
    assert torch.allclose(l, g["full_loss"], rtol=1e-5) and torch.allclose(p, g["full_pred"], rtol=1e-4, atol=1e-5)
    # known answers
    assert torch.equal(o.block_causal_mask(6, 2), g["mask_6_2"])
    assert torch.allclose(torch.view_as_real(o.rope_cache(8, 5)), g["rope_8_5"])
    assert torch.allclose(o.apply_rope(g["rope_in"], o.rope_cache(8, 5)), g["rope_out"])


def test_simple_mae_oracle_matches_reference():
    from oracle import brainformer_ref as o
    g = load("simple_mae_small.pt")
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in g["state"].items()}
    loss, _ = o.simple_mae_forward(sd, g["x"], g["enc_config"], g["mae_config"], g["masked"], g["unmasked"])
    assert torch.allclose(loss, g["loss"], rtol=1e-5), (loss, g["loss"])
    loss.backward()
    for n, ref in g["grads"].items():
        assert torch.allclose(sd[n].grad, ref, rtol=1e-4, atol=1e-6), n


@pytest.mark.parametrize("cosine", [False, True])
def test_vq_oracle_properties(cosine):
    """The quantiser oracle is unpinned (the library is not in the reference tree); check the properties its
    published algorithm guarantees: nearest code, idempotence, STE value, loss definition, EMA bookkeeping."""
    from oracle.vector_quantize_ref import VectorQuantizeRef
    torch.manual_seed(0)
    K, D = 32, 16
    vq = VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine).train()
    x = torch.randn(2, 40, D, requires_grad=True)
    before = vq._codebook.embed.clone()
    q, ind, loss = vq(x)
    xs = torch.nn.functional.normalize(x.detach(), dim=-1) if cosine else x.detach()
    brute = (xs.reshape(-1, D) @ before[0].t()).argmax(-1) if cosine else torch.cdist(xs.reshape(-1, D), before[0]).argmin(-1)
    assert torch.equal(ind.reshape(-1), brute)
    assert ind.dtype == torch.int64 and tuple(loss.shape) == (1,)
    assert torch.allclose(q.detach(), before[0][ind], atol=1e-6)                       # STE value == codeword
    assert torch.allclose(loss, 0.25 * ((before[0][ind] - xs) ** 2).mean().reshape(1), rtol=1e-5)
    bins = torch.bincount(ind.reshape(-1), minlength=K).float()
    assert torch.allclose(vq._codebook.cluster_size[0], 0.2 * bins, atol=1e-6)          # lerp from zeros, decay 0.8
    vq.eval()
    q2, ind2, l2 = vq(vq._codebook.embed[0][None])
    assert torch.equal(ind2.reshape(-1), torch.arange(K)) and float(l2) == 0.0           # idempotence, zero eval loss


def _ddp_worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle.vector_quantize_ref import VectorQuantizeRef
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(3)
    K, D, N = 16, 8, 64
    X = torch.randn(N, D)
    C = torch.randn(K, D)

    def ar(t):
        dist.all_reduce(t)

    vq = VectorQuantizeRef(dim=D, codebook_size=K, use_cosine_sim=False, all_reduce_fn=ar).train()
    vq._codebook.embed.copy_(C[None]); vq._codebook.embed_avg.copy_(C[None]); vq._codebook.cluster_size.fill_(1.0)
    shard = X[rank * (N // world):(rank + 1) * (N // world)]
    vq(shard[None])
    if rank == 0:
        q.put({k: v.clone() for k, v in vq.state_dict().items()})
    dist.destroy_process_group()


def test_vq_oracle_two_rank_allreduce_equals_single_rank():
    """all-reduced (bins, embed_sum) over 2 ranks == the 1-rank run on the concatenated batch."""
    import torch.multiprocessing as mp
    from oracle.vector_quantize_ref import VectorQuantizeRef
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    torch.manual_seed(3)
    K, D, N = 16, 8, 64
    X = torch.randn(N, D)
    C = torch.randn(K, D)
    vq = VectorQuantizeRef(dim=D, codebook_size=K, use_cosine_sim=False).train()
    vq._codebook.embed.copy_(C[None]); vq._codebook.embed_avg.copy_(C[None]); vq._codebook.cluster_size.fill_(1.0)
    vq(X[None])
    for k, v in vq.state_dict().items():
        assert torch.allclose(got[k], v, rtol=1e-5, atol=1e-6), k


def _encodec_lineage_codebook():
    """EuclideanCodebook of the encodec lineage as shipped inside the installed vllm wheel
    (vllm/model_executor/models/mimo_audio.py, "Vector quantization (from MiMo-Audio-Tokenizer)"): NOT the reference's
    dependency, but an independent implementation of the same published algorithm (EMA cluster sizes, EMA embedding
    sums, Laplace smoothing, embed = embed_avg / smoothed size) -- SURVEY.md section 8c names it as the sibling-lineage
    cross-check for the unpinned quantiser oracle.  Only the self-contained VQ section of the file is executed."""
    import importlib.util
    import typing as tp
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F
    einops = pytest.importorskip("einops")
    spec = importlib.util.find_spec("vllm")
    if spec is None or not spec.submodule_search_locations:
        pytest.skip("vllm is not installed")
    path = os.path.join(list(spec.submodule_search_locations)[0], "model_executor", "models", "mimo_audio.py")
    if not os.path.exists(path):
        pytest.skip("this vllm build has no mimo_audio.py")
    src = open(path).read()
    try:
        a = src.index("def _vq_default")
        a_end = src.index("\nclass ", a)
        b = src.index("class EuclideanCodebook")
        b_end = src.index("\nclass ", b + 10)
    except ValueError:
        pytest.skip("mimo_audio.py no longer has the expected VQ section")
    ns = dict(torch=torch, nn=nn, F=F, dist=dist, tp=tp, rearrange=einops.rearrange, repeat=einops.repeat)
    exec(compile(src[a:a_end] + "\n" + src[b:b_end], path, "exec"), ns)
    return ns["EuclideanCodebook"]


def test_vq_oracle_against_encodec_lineage_codebook():
    """Second anchor for the unpinned oracle: three training steps of the oracle's Euclidean codebook and of the
    encodec-lineage one from the same state on the same batches give the same indices, quantised rows and EMA state
    (threshold 0: the two lineages differ in WHEN dead codes are re-drawn, not in the EMA arithmetic)."""
    from oracle.vector_quantize_ref import VectorQuantizeRef
    Codebook = _encodec_lineage_codebook()
    K, D, N = 48, 16, 400
    g = torch.Generator().manual_seed(0)
    C = torch.randn(K, D, generator=g)
    ref = VectorQuantizeRef(dim=D, codebook_size=K, commitment_weight=0.25, decay=0.8, eps=1e-5, threshold_ema_dead_code=0).train()
    ref._codebook.embed.copy_(C[None]); ref._codebook.embed_avg.copy_(C[None]); ref._codebook.cluster_size.fill_(1.0)
    ref._codebook.initted.fill_(1.0)
    other = Codebook(dim=D, codebook_size=K, kmeans_init=False, decay=0.8, epsilon=1e-5, threshold_ema_dead_code=0).train()
    other.embed.copy_(C); other.embed_avg.copy_(C); other.cluster_size.fill_(1.0)
    for step in range(3):
        x = C[torch.randint(0, K, (N,), generator=g)] + 0.4 * torch.randn(N, D, generator=g)
        q_ref, i_ref, _ = ref(x[None])
        q_o, i_o = other(x)
        assert torch.equal(i_ref.view(-1), i_o.view(-1)), step
        # (the oracle returns the straight-through value x + (q - x).detach(): the same numbers up to one rounding)
        assert torch.allclose(q_ref.view(N, D), q_o, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ref._codebook.cluster_size[0], other.cluster_size, rtol=1e-5, atol=1e-6), step
        assert torch.allclose(ref._codebook.embed_avg[0], other.embed_avg, rtol=1e-5, atol=1e-5), step
        assert torch.allclose(ref._codebook.embed[0], other.embed, rtol=1e-4, atol=1e-5), step
