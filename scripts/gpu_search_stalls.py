"""Where does fk::vq_search_kernel wait?  Runs the stall-accounting instantiation (fk_vq_search_profile) and prints the
per-CTA mean / max of every counter as cycles and as a share of the CTA's lifetime.  Diagnosis only: the profiled
instantiation is slower than the product kernel (clock64 reads around every barrier wait)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import vector_quantize as fvq
from frankenstein_b200._lib import lib, ptr, stream, check

NAMES = ["total", "setup", "prod_wait_empty", "prod_wait_x_empty", "iss0_wait_x_full", "iss1_wait_x_full",
         "iss0_wait_tmem_empty", "iss1_wait_tmem_empty", "iss0_wait_full", "iss1_wait_full", "iss0_loop", "iss1_loop",
         "epi_wait_tmem_full", "epi_flush", "epi_named_barrier", "first_acc_ready",
         "iss0_pre", "iss1_pre", "iss0_mma_issue", "iss1_mma_issue", "iss0_commit", "iss1_commit"]


def main():
    shapes = [(16384, 8192, 256), (16384, 65536, 256)]
    out = {}
    for N, K, D in shapes:
        X = torch.randn(N, D, device="cuda"); C = torch.randn(K, D, device="cuda")
        _, xb, _ = fvq.prepare_input(X, False)
        cb, c2 = fvq.prepare_codebook(C, False)
        G = torch.cuda.get_device_properties(0).multi_processor_count
        S = lib().fk_vq_search_slots(N, K, G)
        cv = torch.empty(N, S, 4, device="cuda"); ci = torch.empty(N, S, 4, device="cuda", dtype=torch.int32)
        prof = torch.zeros(G, 32, device="cuda", dtype=torch.int64)
        for _ in range(3):
            check(lib().fk_vq_search_profile(ptr(xb), ptr(cb), ptr(c2), N, K, xb.shape[1], 0, ptr(cv), ptr(ci), S, G,
                                             ptr(prof), stream()), "fk_vq_search_profile")
        torch.cuda.synchronize()
        raw = prof.cpu()
        cnt = (raw >> 40).double()                  # number of waits that actually blocked
        p = (raw & ((1 << 40) - 1)).double()
        tot = p[:, 0].mean().item()
        print(f"N={N} K={K} D={D}: CTA lifetime mean {tot:.0f} cycles, max {p[:, 0].max().item():.0f}")
        rec = {}
        for i, n in enumerate(NAMES):
            rec[n] = dict(mean=p[:, i].mean().item(), max=p[:, i].max().item())
            print(f"  {n:24s} mean {p[:, i].mean().item():10.0f}  max {p[:, i].max().item():10.0f}  ({100 * p[:, i].mean().item() / tot:5.1f} % of lifetime)  blocked waits {cnt[:, i].mean().item():7.1f}")
        out[f"{N}x{K}x{D}"] = rec
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/search_stalls.json", "w"), indent=1)


if __name__ == "__main__":
    main()
