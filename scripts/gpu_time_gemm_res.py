import os, sys, json, torch
sys.path.insert(0, "/root/repo")
from frankenstein_b200 import gemm
M=524288
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(M,512,device="cuda",generator=g)).to(torch.bfloat16)
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
for N in (512,1536,4096):
    w=(torch.randn(N,512,device="cuda",generator=g)*0.04).to(torch.bfloat16)
    ms=t(lambda: gemm.gemm_nt(x,w))
    print(os.environ.get("FK_GEMM_DBG","0"), N, round(ms,4), "ms", round(2.0*M*N*512/ms/1e9,1), "TF/s", flush=True)
