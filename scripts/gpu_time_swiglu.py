import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import ops
M, H = 524288, 2048
h = torch.randn(M, 2 * H, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
g = torch.randn(M, H, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    y = ops.swiglu(h); y.backward(g); h.grad = None
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); y = ops.swiglu(h); e[1].record(); y.backward(g); e[2].record(); torch.cuda.synchronize()
print("swiglu fwd %.3f ms (%.2f TB/s)  bwd %.3f ms (%.2f TB/s)" % (e[0].elapsed_time(e[1]), M*H*6/e[0].elapsed_time(e[1])/1e9, e[1].elapsed_time(e[2]), M*H*10/e[1].elapsed_time(e[2])/1e9))
