"""torchrun --nproc-per-node 2 scripts/gpu_ddp_check.py : data-parallel semantics of the quantiser on real GPUs.
(1) 2-rank sharded step with the packed EMA all-reduce == 1-rank step on the concatenated batch;
(2) with dead-code reset enabled, all ranks end with bit-identical codebooks (candidates travel in the same all-reduce)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from frankenstein_b200.vector_quantize import VectorQuantize


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
        for name in ("embed", "embed_avg", "cluster_size"):
            da = (getattr(a._codebook, name) - getattr(b._codebook, name)).abs().max().item()
            sc = getattr(b._codebook, name).abs().max().item()
            if da > 1e-4 * sc + 1e-6:
                ok = False
                print(f"rank {rank} cosine={cosine} {name}: sharded vs single max diff {da} (scale {sc})")
        c = fresh(2, True)
        for _ in range(3):
            c(shard[None])
        mine = c._codebook.embed.clone()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        if not torch.equal(mine, ref):
            ok = False
            print(f"rank {rank} cosine={cosine}: codebooks differ across ranks after dead-code reset")
        nexp = int(c.last_n_expired.item())
        if rank == 0:
            print(f"cosine={cosine}: sharded==single ok, ranks identical, expired codes last step = {nexp}")
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DDP CHECK", "PASSED" if int(t.item()) == 1 else "FAILED")
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
