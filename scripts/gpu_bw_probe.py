import torch
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/reps
n = 2 << 30
x = torch.empty(n, dtype=torch.uint8, device="cuda"); y = torch.empty(n, dtype=torch.uint8, device="cuda")
ms = t(lambda: x.zero_()); print("memset 2GiB", round(n/ms/1e9,2), "TB/s write")
xf = x.view(torch.float32)
ms = t(lambda: xf.fill_(1.5)); print("fill 2GiB", round(n/ms/1e9,2), "TB/s write")
ms = t(lambda: y.copy_(x)); print("copy 2GiB", round(2*n/ms/1e9,2), "TB/s r+w")
ms = t(lambda: xf.sum()); print("read-reduce 2GiB", round(n/ms/1e9,2), "TB/s read")
# strided tile-like write: [M, 1536] bf16 written in column blocks of 128 B
M=524288
c = torch.empty(M,1536,dtype=torch.bfloat16,device="cuda")
src = torch.randn(M,64,device="cuda").to(torch.bfloat16)
def colblocks():
    for j in range(0,1536,64): c[:, j:j+64].copy_(src)
ms = t(colblocks, 3); print("column-block writes (128B rows) 1.6GB", round(c.numel()*2/ms/1e9,2), "TB/s write (+small read)")
