"""Where do the tcgen05 attention-backward CTAs spend their cycles?  Runs fk_attn_backward_tc with the stall-accounting
instantiation (fk_attn_backward_tc_profile, selected through ops._BWD_PROFILE) at the cfg-2 shape and prints per-CTA means.  Diagnosis only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import ops
from frankenstein_b200._lib import lib, ptr, check

NAMES = ["lifetime", "setup", "tiles", "first_scores_ready@", "wg0_wait_sdp_full", "(unused)", "wg0_named_barrier",
         "wg0_compute", "wg0_last_p_ready@", "acc_complete@", "stores_done@", "score_iss_wait_st_full",
         "score_iss_wait_stage_free", "acc_iss_wait_p_ready", "producer_wait_st_empty", "wg0_first_tmem_ld_wait",
         "wg0_later_tmem_ld_waits", "wg0_tmem_st_wait"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    S, H = 4096, 16
    dev = torch.device("cuda")
    qkv = torch.randn(B, S, 3 * H * 32, device=dev, dtype=torch.bfloat16)
    mask = ops.LabelMask.block_causal(B, S, 256, dev)
    w = torch.randn(B, S, H * 32, device=dev, dtype=torch.bfloat16)
    n_cta = B * H * (S // 128)
    for parts, name in ((2, "dK/dV kernel"), (4, "dQ kernel")):
        prof = torch.zeros(n_cta * 24 + 6 * 128, device=dev, dtype=torch.int64)
        ops._BWD_PARTS = (parts,)
        for it in range(2):                          # warm-up without the buffer, then ONE profiled launch
            x = qkv.clone().requires_grad_(True)
            out = ops.attention_qkv(x * 1.0, H, None, mask)
            if it == 1:
                ops._BWD_PROFILE = (prof, 1)
            out.backward(w)
        ops._BWD_PARTS = (2, 4)
        torch.cuda.synchronize()
        ops._BWD_PROFILE = None
        tr = prof[n_cta * 24:].cpu().view(6, 128)
        p = prof[:n_cta * 24].view(n_cta, 24).double().cpu()
        p = p[p[:, 2] > 0]
        life = p[:, 0].mean().item()
        T = p[:, 2].mean().item()
        print(f"{name}: {p.shape[0]} CTAs, mean tiles {T:.1f}, mean lifetime {life:.0f} cycles = {life / T:.0f} per tile")
        for i, n in enumerate(NAMES):
            print(f"  {n:26s} mean {p[:, i].mean().item():9.0f}  ({100 * p[:, i].mean().item() / life:5.1f} % of lifetime)")
        # per-tile event trace of the first item of CTA 0: 0 acc issuer saw p_ready, 1 acc MMAs issued, 2 score issuer saw
        # stage_free, 3 score MMAs issued, 4 warpgroup saw sdp_full, 5 warpgroup arrived on p_ready
        nt = int((tr[5] > 0).sum().item())
        if nt > 12 and bool((tr[:, :nt] > 0).all()):
            sl = slice(6, nt - 3)
            a0, a1, s2, s3, w4, w5 = [tr[i].double() for i in range(6)]
            hop1 = (a0[sl] - w5[sl]).mean().item()                       # p_ready arrive -> acc issuer awake
            iss1 = (a1[sl] - a0[sl]).mean().item()                       # acc issue
            j3 = slice(9, nt)                                            # tile j+3 reuses the stage of tile j
            j0_ = slice(6, nt - 3)
            hop2 = (s2[j3] - a1[j0_]).mean().item()                     # acc MMAs exec + stage_free -> score issuer awake
            iss2 = (s3[j3] - s2[j3]).mean().item()
            hop3 = (w4[j3] - s3[j3]).mean().item()                       # score MMAs exec + sdp_full -> warpgroup awake
            comp = (w5[sl] - w4[sl]).mean().item()
            print(f"  stage turnaround trace (tiles 6..{nt - 4}): p_ready->acc issuer {hop1:.0f}, acc issue {iss1:.0f}, "
                  f"acc exec + stage_free->score issuer {hop2:.0f}, score issue {iss2:.0f}, score exec + sdp_full->warpgroup {hop3:.0f}, "
                  f"compute (sdp_full seen -> p_ready) {comp:.0f} cycles")
        steady = (p[:, 8] - p[:, 3]).mean().item()
        print(f"  steady state (first scores -> last P ready): {steady:.0f} cycles = {steady / T:.0f} per tile; "
              f"prologue {p[:, 3].mean().item():.0f}, tail {(p[:, 0] - p[:, 8]).mean().item():.0f}")
    # ---- light mode: the product code path plus 2 timestamps per CTA -> true lifetimes and the idle gap between
    #      consecutive CTAs on one SM (the kernel runs one CTA per SM)
    for parts, name in ((2, "dK/dV kernel"), (4, "dQ kernel")):
        prof = torch.zeros(n_cta, 24, device=dev, dtype=torch.int64)
        ops._BWD_PROFILE = (prof, 2)
        ops._BWD_PARTS = (parts,)
        x = qkv.clone().requires_grad_(True)
        out = ops.attention_qkv(x * 1.0, H, None, mask)
        out.backward(w)
        ops._BWD_PARTS = (2, 4)
        torch.cuda.synchronize()
        ops._BWD_PROFILE = None
        p = prof.cpu()
        life, T, g0, g1, sm = p[:, 0].double(), p[:, 2].double(), p[:, 20], p[:, 21], p[:, 22]
        gaps, busy = [], []
        for s_id in sm.unique().tolist():
            idx = (sm == s_id).nonzero().flatten()
            order = idx[g0[idx].argsort()]
            st, en = g0[order], g1[order]
            if len(order) > 1:
                gaps.append((st[1:] - en[:-1]).double())
            busy.append((en - st).double())
        gaps = torch.cat(gaps); busy = torch.cat(busy)
        span = (g1.max() - g0.min()).item()
        print(f"{name} (light): kernel span {span / 1e3:.1f} us; CTA lifetime mean {life.mean().item():.0f} cycles "
              f"({busy.mean().item():.0f} ns) = {life.sum().item() / T.sum().item():.0f} cycles per tile; "
              f"gap between CTAs on an SM: mean {gaps.mean().item():.0f} ns, median {gaps.median().item():.0f} ns; "
              f"SM busy share {busy.sum().item() / (span * len(sm.unique())):.3f}")


if __name__ == "__main__":
    main()
