"""One launch of each A-resident GEMM variant at the cfg-4 shape (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from frankenstein_b200 import gemm
M = 524288
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, 512, device="cuda", generator=g).to(torch.bfloat16)
w = (torch.randn(1536, 512, device="cuda", generator=g) * 0.04).to(torch.bfloat16)
w13 = (torch.randn(4096, 512, device="cuda", generator=g) * 0.04).to(torch.bfloat16)
for _ in range(2):
    gemm.gemm_nt(x, w)
    h13, gated = gemm.gemm_nt(x, w13, None, gemm.EPI_SWIGLU)
torch.cuda.synchronize()
