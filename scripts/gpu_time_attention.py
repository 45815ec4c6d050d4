"""Time the attention kernels at the cfg-2 shape (per-trial S=4096, 16 heads x 32, block-causal E=256)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import ops

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    S, H = 4096, 16
    dev = torch.device("cuda")
    qkv = torch.randn(B, S, 3 * H * 32, device=dev, dtype=torch.bfloat16)
    mask = ops.LabelMask.block_causal(B, S, 256, dev)
    w = torch.randn(B, S, H * 32, device=dev, dtype=torch.bfloat16)
    dens = 17 / 32.0
    fl = 4.0 * B * H * S * S * 32 * dens
    res = {}
    for it in range(4):
        x = qkv.clone().requires_grad_(True)
        y = x * 1.0
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(); out = ops.attention_qkv(y, H, None, mask); e[1].record()
        e[2].record(); out.backward(w); e[3].record()
        torch.cuda.synchronize()
        res = dict(B=B, fwd_ms=e[0].elapsed_time(e[1]), bwd_ms=e[2].elapsed_time(e[3]))
    res["fwd_tflops"] = fl / res["fwd_ms"] / 1e9
    res["bwd_tflops"] = 2.5 * fl / res["bwd_ms"] / 1e9
    print(json.dumps(res))

if __name__ == "__main__":
    main()
