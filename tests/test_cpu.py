"""CPU: the C-ABI library loads and exports every symbol of include/fk_b200.h, the host-side mirrors keep the
reference's state-dict keys / parameter counts / mask + rope conventions, and the product path refuses CPU tensors."""
import ctypes
import os

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol():
    from frankenstein_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 30
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(L, name), f"{name} declared in include/fk_b200.h but not exported"
    L2 = _lib.lib()
    assert L2.fk_abi_version() == 2 and L2.fk_target_sm() == 100
    assert L2.fk_last_error() is not None


def test_host_helpers_without_gpu():
    from frankenstein_b200 import _lib
    L = _lib.lib()
    # slots: 2 column halves x number of CTAs that can touch one row block
    assert L.fk_vq_search_slots(16384, 8192, 148) == 8
    assert L.fk_vq_search_slots(256, 128, 148) == 2
    assert L.fk_vq_finish_partials(17) == 3 and L.fk_masked_l1_partials(16) == 2
    assert L.fk_norm_backward_grid() == 148
    if not torch.cuda.is_available():
        assert L.fk_device_ok() == 0
    # argument validation happens before any CUDA call
    assert L.fk_vq_search(None, None, None, 0, 0, 64, 0, None, None, 1, 0, None) == -1
    assert b"fk_vq_search" in L.fk_last_error()


def test_dead_code_layout():
    from frankenstein_b200.vector_quantize import dead_code_layout
    assert dead_code_layout(8192, 16384, 1) == (8192, 8192)
    assert dead_code_layout(8192, 16384, 8) == (1024, 8192)
    assert dead_code_layout(512, 100, 2) == (100, 200)
    assert dead_code_layout(10, 3, 4) == (2, 8)


def test_cpu_tensors_are_rejected_loudly():
    from frankenstein_b200._lib import FkError
    from frankenstein_b200.vector_quantize import VectorQuantize
    from frankenstein_b200 import brainformer as bf
    with pytest.raises(FkError):
        VectorQuantize(dim=64, codebook_size=16)(torch.zeros(1, 4, 64))
    enc = bf.Encoder(bf.MAEConfig(window_size=64, n_electrodes=8, patch_size=8, dim=64, n_layers=1, head_dim=32, hidden_dim=64,
                                  n_heads=2, n_kv_heads=2))
    with pytest.raises(Exception):
        enc(torch.zeros(1, 64, 8))


def test_state_dict_keys_match_reference_checkpoints():
    from frankenstein_b200 import brainformer as bf
    from frankenstein_b200 import simple_mae as sm
    from frankenstein_b200.vq_brain import SoundStream
    g = torch.load(os.path.join(GOLD, "soundstream_cosine.pt"), weights_only=False)
    m = SoundStream(**g["config"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: tuple(v.shape) for k, v in g["state_dict"].items()}
    m.load_state_dict(g["state_dict"], strict=True)
    b = torch.load(os.path.join(GOLD, "brainformer_small.pt"), weights_only=False)
    enc = bf.Encoder(bf.MAEConfig(**b["enc_config"]))
    enc.load_state_dict(b["enc_state"], strict=True)
    mae = bf.MAE(bf.MAEConfig(**b["enc_config"]))
    mae.load_state_dict(b["mae_state"], strict=True)
    full = bf.BrainFormer(bf.Config(encoder=bf.MAEConfig(**b["enc_config"]), **b["per_config"]))
    full.load_state_dict(b["full_state"], strict=True)
    s = torch.load(os.path.join(GOLD, "simple_mae_small.pt"), weights_only=False)
    smae = sm.SimpleMAE(sm.SimpleEncoderConfig(**s["enc_config"]), sm.SimpleMAEConfig(**s["mae_config"]))
    smae.load_state_dict(s["state"], strict=True)


def test_known_answers():
    """5.61M / 4.27M / 6.32M parameters, [6144,6144] mask, mask(6,2) and rope micro-vectors."""
    from frankenstein_b200 import brainformer as bf
    from frankenstein_b200.vq_brain import SoundStream
    b = torch.load(os.path.join(GOLD, "brainformer_small.pt"), weights_only=False)
    ss = SoundStream(C=256, D=64, codebook_size=1024, n_electrodes=512)
    assert round(sum(p.numel() for p in ss.parameters()) / 1e6, 2) == 5.61
    mc = bf.MAEConfig(window_size=768, patch_size=32)
    enc = bf.Encoder(mc)
    assert enc.get_num_params() == b["params_encoder_768_32"] and round(enc.get_num_params() / 1e6, 2) == 4.27
    assert tuple(enc.attn_mask.shape) == (6144, 6144)
    full = bf.BrainFormer(bf.Config(encoder=mc, n_output_tokens=32, output_dim=768))
    assert round(full.get_num_params() / 1e6, 2) == 6.32
    assert torch.equal(bf.build_advanced_causal_mask(6, 2), b["mask_6_2"])
    assert torch.allclose(torch.view_as_real(bf.build_complex_rope_cache(8, 5, 10000)), b["rope_8_5"])
    assert torch.allclose(bf.apply_rope(b["rope_in"], bf.build_complex_rope_cache(8, 5, 10000)), b["rope_out"])
    x = torch.arange(2 * 16 * 4, dtype=torch.float32).view(2, 16, 4)
    small = bf.Encoder(bf.MAEConfig(window_size=16, n_electrodes=4, patch_size=8, dim=32, n_layers=1, head_dim=32, hidden_dim=32,
                                    n_heads=1, n_kv_heads=1))
    from oracle.brainformer_ref import to_patches
    assert torch.equal(small.to_patches(x), to_patches(x, 8))


def test_label_rule_reproduces_the_reference_masks():
    """The kernels never see a mask tensor: key j is visible to query i iff kid[j] <= qid[i] (ops.LabelMask).  Check on the
    CPU that this rule gives exactly the three dense masks the reference builds: the block-causal mask
    (brainformer.py:93-111, pinned by the golden mask(6, 2)), the MAE sub-matrix gathered at the kept positions
    (brainformer.py:392-413) and the simple_mae padding mask (simple_mae:349-352)."""
    from frankenstein_b200 import brainformer as bf
    from oracle.brainformer_ref import block_causal_mask
    b = torch.load(os.path.join(GOLD, "brainformer_small.pt"), weights_only=False)
    # block causal: labels = position // tokens_per_time_step
    for S, E in ((12, 2), (48, 16), (96, 32)):
        ids = torch.arange(S) // E
        rule = ids[None, :] <= ids[:, None]
        assert torch.equal(rule, bf.build_advanced_causal_mask(S, E).bool())
        assert torch.equal(rule, block_causal_mask(S, E).bool())
    ids = torch.arange(6) // 2                                  # the reference's own 6-token, 2-per-step example
    assert torch.equal((ids[None, :] <= ids[:, None]), b["mask_6_2"].bool())
    # MAE: the reference gathers mask[idx][:, idx] per sample; the label rule gathers the labels instead
    g = torch.Generator().manual_seed(0)
    S, E, keep = 96, 16, 24
    full = bf.build_advanced_causal_mask(S, E).bool()
    for _ in range(4):
        idx = torch.sort(torch.randperm(S, generator=g)[:keep])[0]
        sub = full[idx][:, idx]
        lab = idx // E
        assert torch.equal(lab[None, :] <= lab[:, None], sub)
    # padding: attend iff neither token is padded; kid = pad ? INT_MAX : 0, qid = pad ? -1 : 0
    pad = torch.rand(3, 40, generator=g) < 0.3
    big = torch.iinfo(torch.int32).max
    kid = torch.where(pad, big, 0)
    qid = torch.where(pad, -1, 0)
    rule = kid[:, None, :] <= qid[:, :, None]
    ref = (~pad)[:, :, None] & (~pad)[:, None, :]
    assert torch.equal(rule, ref)


def _packed_ema_worker(rank, world, port, q):
    """One rank of the packed-buffer protocol of VectorQuantize._ema_step, with the rank-local statistics formed by plain
    torch on the CPU (the kernels need a B200): [K*D embed_sum | K bins | R*D dead-code candidate rows]."""
    import torch.distributed as dist
    from frankenstein_b200.vector_quantize import dead_code_layout
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    torch.manual_seed(5)
    K, D, N = 12, 8, 40                                  # N rows per rank
    X = torch.randn(world * N, D)
    ind = torch.randint(0, K, (world * N,))
    xr, ir = X[rank * N:(rank + 1) * N], ind[rank * N:(rank + 1) * N]
    per_rank, R = dead_code_layout(K, N, world)
    stats = torch.zeros(K * D + K + R * D)
    stats[:K * D].view(K, D).index_add_(0, ir, xr)
    stats[K * D:K * D + K].index_add_(0, ir, torch.ones(N))
    rows = (torch.arange(per_rank) * 7 + rank) % N       # this rank's candidate rows (any deterministic choice)
    tail = stats[K * D + K:].view(R, D)
    tail[rank * per_rank:(rank + 1) * per_rank] = xr[rows]     # own slice only; the SUM gathers the slices
    dist.all_reduce(stats)
    q.put(stats.clone() if rank == 0 else None)
    # every rank must hold the identical buffer (codebooks stay bit-identical without buffer broadcasts)
    gathered = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(gathered, stats)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    dist.destroy_process_group()


def test_packed_ema_allreduce_two_ranks_gloo():
    """SURVEY 8e on the CPU: ONE all-reduce of the packed buffer carries embed_sum, bins and the dead-code candidates
    (each rank writes only its slice of the tail); the result equals the single-process statistics of the concatenated
    batch and the tail is the concatenation of the ranks' rows."""
    import torch.multiprocessing as mp
    from frankenstein_b200.vector_quantize import dead_code_layout
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 300
    procs = [ctx.Process(target=_packed_ema_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    got = [q.get(timeout=180) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    stats = next(g for g in got if g is not None)
    torch.manual_seed(5)
    K, D, N = 12, 8, 40
    X = torch.randn(world * N, D)
    ind = torch.randint(0, K, (world * N,))
    per_rank, R = dead_code_layout(K, N, world)
    assert R == per_rank * world and per_rank >= 1
    es = torch.zeros(K, D).index_add_(0, ind, X)
    bins = torch.zeros(K).index_add_(0, ind, torch.ones(world * N))
    assert torch.allclose(stats[:K * D].view(K, D), es, atol=1e-6)
    assert torch.equal(stats[K * D:K * D + K], bins)
    tail = stats[K * D + K:].view(R, D)
    for r in range(world):
        rows = (torch.arange(per_rank) * 7 + r) % N
        assert torch.equal(tail[r * per_rank:(r + 1) * per_rank], X[r * N:(r + 1) * N][rows])
