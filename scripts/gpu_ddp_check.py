"""torchrun --nproc-per-node 2 scripts/gpu_ddp_check.py : data-parallel semantics on real GPUs (NCCL).
(1) quantiser: 2-rank sharded step with the packed EMA all-reduce == 1-rank step on the concatenated batch;
(2) with dead-code reset enabled, all ranks end with bit-identical codebooks (candidates travel in the same all-reduce);
(3) full joint model (SoundStream + BrainFormer) under DistributedDataParallel with bf16 gradient buckets: the averaged
    gradients of the sharded batch == the gradients of one rank on the whole batch, and the codebooks after the step agree."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from frankenstein_b200.vector_quantize import VectorQuantize


class SmallJoint(torch.nn.Module):
    """the bench's joint step at test size: loss = SoundStream(x)[0] + BrainFormer(x[..., :E], t)[0]"""

    def __init__(self):
        super().__init__()
        from frankenstein_b200.brainformer import BrainFormer, Config, MAEConfig
        from frankenstein_b200.vq_brain import SoundStream
        enc = MAEConfig(window_size=128, n_electrodes=64, patch_size=32, dim=128, n_layers=2, head_dim=32, hidden_dim=256, n_heads=4,
                        n_kv_heads=4)
        self.vqvae = SoundStream(C=32, D=64, codebook_size=128, n_electrodes=128, use_cosine_sim=False)
        self.brainformer = BrainFormer(Config(encoder=enc, n_output_tokens=8, output_dim=32, dim=128, n_layers=1, head_dim=16,
                                              hidden_dim=256, n_heads=4, n_kv_heads=4))

    def forward(self, x, t):
        return (self.vqvae(x)[0] + self.brainformer(x[..., :64].contiguous(), t)[0]).sum()


def joint_model_check(rank, world, lr):
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    dev = torch.device("cuda", lr)
    g = torch.Generator().manual_seed(5)
    Bfull = 4 * world
    x = torch.randn(Bfull, 128, 128, generator=g).to(dev)          # no padded rows: every loss term is a plain batch mean
    t = torch.randn(Bfull, 8, 32, generator=g).to(dev)
    cb = torch.randn(128, 64, generator=g).to(dev) * 0.3

    def fresh():
        torch.manual_seed(11)
        m = SmallJoint().to(dev).train()
        q = m.vqvae.quantizer
        q._codebook.embed.copy_(cb[None]); q._codebook.embed_avg.copy_(cb[None]); q._codebook.cluster_size.fill_(3.0)
        q._codebook.initted.fill_(1.0); q._mark_dirty(); q._kmeans_initted_host = True
        q.threshold_ema_dead_code = 0
        return m

    single = fresh()
    single.vqvae.quantizer.sync_codebook = False
    with torch.autocast("cuda", dtype=torch.bfloat16):
        single(x, t).backward()
    sharded = fresh()
    ddp = torch.nn.parallel.DistributedDataParallel(sharded, device_ids=[lr], broadcast_buffers=False, gradient_as_bucket_view=True)
    ddp.register_comm_hook(None, default_hooks.bf16_compress_hook)
    n = Bfull // world
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ddp(x[rank * n:(rank + 1) * n], t[rank * n:(rank + 1) * n]).backward()
    torch.cuda.synchronize()
    ok, worst, worst_name = True, 0.0, ""
    ps, pd = dict(single.named_parameters()), dict(sharded.named_parameters())
    for name, p in ps.items():
        if p.grad is None:
            continue
        a, b = pd[name].grad.float(), p.grad.float()
        err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
        if err > worst:
            worst, worst_name = err, name
        if err > 4e-2:
            ok = False
            print(f"rank {rank}: gradient of {name} differs: rel {err:.3e}")
    sharded.vqvae.quantizer._wait_pending(); single.vqvae.quantizer._wait_pending()
    for name in ("embed", "embed_avg", "cluster_size"):
        a = getattr(sharded.vqvae.quantizer._codebook, name)
        b = getattr(single.vqvae.quantizer._codebook, name)
        d = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
        if d > 1e-4:
            ok = False
            print(f"rank {rank}: codebook buffer {name} differs after the step: rel {d:.3e}")
    if rank == 0:
        print(f"joint model, {world} ranks x {n} trials vs 1 rank x {Bfull}: {len(ps)} parameter gradients agree "
              f"(worst relative difference {worst:.2e} at {worst_name}; bf16 buckets), codebooks agree")
    return ok


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    K, D, N = 256, 64, 4096
    g = torch.Generator().manual_seed(0)
    X = torch.randn(N, D, generator=g).cuda()
    C = torch.randn(K, D, generator=g).cuda()
    ok = True
    for cosine in (False, True):
        def fresh(thr, sync):
            m = VectorQuantize(dim=D, codebook_size=K, commitment_weight=0.25, use_cosine_sim=cosine, threshold_ema_dead_code=thr,
                               sync_codebook=sync).cuda().train()
            c = torch.nn.functional.normalize(C, dim=-1) if cosine else C
            m._codebook.embed.copy_(c[None]); m._codebook.embed_avg.copy_(c[None]); m._codebook.cluster_size.fill_(1.0)
            m._mark_dirty(); m._kmeans_initted_host = True
            return m
        shard = X[rank * (N // world):(rank + 1) * (N // world)]
        a = fresh(0, True)
        for _ in range(3):
            a(shard[None])
        b = fresh(0, False)
        for _ in range(3):
            b(X[None])
        a._wait_pending(); b._wait_pending()
        for name in ("embed", "embed_avg", "cluster_size"):
            da = (getattr(a._codebook, name) - getattr(b._codebook, name)).abs().max().item()
            sc = getattr(b._codebook, name).abs().max().item()
            if da > 1e-4 * sc + 1e-6:
                ok = False
                print(f"rank {rank} cosine={cosine} {name}: sharded vs single max diff {da} (scale {sc})")
        c = fresh(2, True)
        for _ in range(3):
            c(shard[None])
        mine = c.state_dict()["_codebook.embed"].clone()      # (state_dict() waits for the side-stream EMA update of the last step)
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(mine, ref):
            ok = False
            print(f"rank {rank} cosine={cosine}: codebooks differ across ranks after dead-code reset")
        nexp = int(c.last_n_expired.item())
        if rank == 0:
            print(f"cosine={cosine}: sharded==single ok, ranks identical, expired codes last step = {nexp}")
    ok = joint_model_check(rank, world, lr) and ok
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP CHECK", "PASSED" if int(t.item()) == 1 else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
