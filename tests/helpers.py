"""Seeded synthetic inputs shared by the parity tests."""
import torch


def make_vq_problem(N, K, D, seed=0, cosine=False, noise=0.3, planted_ties=0):
    g = torch.Generator().manual_seed(seed)
    C0 = torch.randn(K, D, generator=g)
    if cosine:
        C0 = torch.nn.functional.normalize(C0, dim=-1)
    assign = torch.randint(0, K, (N,), generator=g)
    X = C0[assign] + noise * torch.randn(N, D, generator=g) * (C0.norm(dim=-1).mean() / D ** 0.5 if cosine else 1.0)
    if planted_ties:
        # exact ties: duplicate codewords (lowest index must win) and rows equal to a codeword
        for i in range(planted_ties):
            a, b = 2 * i, K - 1 - 2 * i
            C0[b] = C0[a]
            X[i] = C0[a]
    return X.contiguous(), C0.contiguous()


def random_vq_problem(N, K, D, seed=0):
    """Unstructured Gaussian rows vs Gaussian codes: many near-ties, the hard case for bf16."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(N, D, generator=g), torch.randn(K, D, generator=g)


def synth_trials(B, T=512, C=512, seed=1234):
    """smoothed, per-channel z-scored noise with a zero-padded tail per trial."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T + 4, C, generator=g)
    x = (x[:, 0:T] + x[:, 1:T + 1] + x[:, 2:T + 2] + x[:, 3:T + 3] + x[:, 4:T + 4]) / 5.0
    x = (x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True)
    pad = torch.randint(0, T // 4 + 1, (B,), generator=g)
    for b in range(B):
        if pad[b] > 0:
            x[b, T - int(pad[b]):] = 0
    return x.contiguous()


def index_agreement(ind, ref_ind, X, C, cosine):
    """(fraction equal, worst relative score gap over the mismatches) measured with the oracle's fp32 scores."""
    ind = ind.reshape(-1).cpu()
    ref_ind = ref_ind.reshape(-1).cpu()
    mism = (ind != ref_ind).nonzero().flatten()
    agree = 1.0 - mism.numel() / ind.numel()
    worst = 0.0
    if mism.numel():
        x = X[mism].double()
        a, b = C[ind[mism]].double(), C[ref_ind[mism]].double()
        if cosine:
            sa, sb = (x * a).sum(-1), (x * b).sum(-1)
            gap = (sb - sa).abs() / sb.abs().clamp_min(1e-12)
        else:
            da, db = (x - a).norm(dim=-1), (x - b).norm(dim=-1)
            gap = (da - db).abs() / db.clamp_min(1e-12)
        worst = float(gap.max())
    return agree, worst
