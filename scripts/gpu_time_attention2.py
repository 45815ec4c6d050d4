"""Per-kernel attention timings (forward, transpose, dK/dV, dQ) for several shapes / masks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import ops, _lib

def run(B, S, H, mask_kind):
    dev = torch.device("cuda")
    qkv = torch.randn(B, S, 3 * H * 32, device=dev, dtype=torch.bfloat16)
    mask = ops.LabelMask.block_causal(B, S, 256, dev) if mask_kind == "block256" else None
    w = torch.randn(B, S, H * 32, device=dev, dtype=torch.bfloat16)
    for it in range(3):
        if it == 2:
            _lib.TIMER.reset(); _lib.TIMER.enabled = True
        x = qkv.clone().requires_grad_(True)
        out = ops.attention_qkv(x * 1.0, H, None, mask)
        out.backward(w)
    _lib.TIMER.enabled = False
    s = _lib.TIMER.summary()
    print(json.dumps(dict(B=B, S=S, H=H, mask=mask_kind, impl=ops.ATTN_BWD_IMPL, **{k: round(v[1], 4) for k, v in s.items()})), flush=True)

if __name__ == "__main__":
    run(16, 4096, 16, "block256")
    run(16, 4096, 16, "none")
    run(16, 1024, 16, "none")
    run(64, 1024, 16, "none")
    run(4, 4096, 16, "none")
