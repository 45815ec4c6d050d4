"""Every kernel of the library timed alone at the cfg-4 shapes (one B200): CUDA events on the launching stream, 2 warm-ups
+ 5 timed launches, a 256 MB write between launches whose inputs do not exceed the L2 by themselves.  One JSON line per
kernel: ms, the algorithmic bytes or flops of the launch and the fraction of the measured peak
(MEASURED_PEAKS.json: hbm_gbs, bf16_tflops = burst figure for a kernel timed alone).

  python scripts/gpu_kernel_table.py            -> JSON lines (profiles/r02_kernel_table.jsonl)
  python scripts/gpu_kernel_table.py --once     -> one launch of each kernel, for `ncu --set full` captures
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

from frankenstein_b200 import data_pipeline as dp
from frankenstein_b200 import gemm, ops, prefix
from frankenstein_b200 import vector_quantize as fvq
from frankenstein_b200._lib import check, lib, ptr, stream
from frankenstein_b200.brainformer import build_complex_rope_cache
from frankenstein_b200.vq_brain import SoundStream, perplexity

ONCE = "--once" in sys.argv
DEV = torch.device("cuda")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6550.0)
TC = PEAKS.get("bf16_tflops", 1650.0)
FLUSH = torch.empty(256 << 20, device=DEV, dtype=torch.uint8)
G = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, dtype=torch.float32, s=1.0):
    return (torch.randn(*shape, device=DEV, generator=G) * s).to(dtype)


def measure(fn, flush):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 0.0
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(5):
        if flush:
            FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / 5


def report(name, fn, bytes_=None, flops=None, flush=True, note=""):
    ms = measure(fn, flush)
    rec = {"kernel": name, "ms": round(ms, 4)}
    if ms > 0 and bytes_ is not None:
        rec.update(bound="hbm", algorithmic_MB=round(bytes_ / 1e6, 2), achieved_GBps=round(bytes_ / ms / 1e6, 1),
                   frac_of_hbm_peak=round(bytes_ / ms / 1e6 / HBM, 3))
    if ms > 0 and flops is not None:
        rec.update(bound="tensor", algorithmic_GF=round(flops / 1e9, 2), achieved_TFps=round(flops / ms / 1e9, 1),
                   frac_of_burst_peak=round(flops / ms / 1e9 / TC, 3))
    if note:
        rec["note"] = note
    print(json.dumps(rec), flush=True)


def main():
    M, D, HID = int(os.environ.get("FK_TABLE_M", 524288)), 512, 2048      # 128 trials x 4096 tokens (FK_TABLE_M=65536: 16 trials, for ncu replays)
    # ---- residual stream ----
    x = rnd(M, D)
    delta = rnd(M, D, dtype=torch.bfloat16)
    w, b = torch.ones(D, device=DEV), torch.zeros(D, device=DEV)
    xg = x.clone().requires_grad_(True)
    dg = delta.clone().requires_grad_(True)
    h, y = ops.add_layer_norm(xg, dg, w.requires_grad_(True), b.requires_grad_(True))
    gy, gh = rnd(M, D, dtype=torch.bfloat16), rnd(M, D)
    report("add_norm_fwd (x + delta -> x', LayerNorm(x') bf16)", lambda: ops.add_layer_norm(x, delta, w, b), bytes_=M * D * 12, flush=False)
    report("add_norm_bwd (dy bf16, dx' fp32 -> dx fp32 + bf16)", lambda: torch.autograd.grad((h, y), (xg, dg), (gh, gy), retain_graph=True),
           bytes_=M * D * 16, flush=False, note="includes the two torch partial-sum reductions of dweight / dbias")
    del x, delta, xg, dg, h, y, gy, gh
    # ---- SwiGLU derivative (block-interleaved h13) ----
    h13 = rnd(M, 2 * HID, dtype=torch.bfloat16)
    dgt = rnd(M, HID, dtype=torch.bfloat16)
    dh13 = torch.empty_like(h13)
    report("swiglu_bwd (blocked h13)", lambda: check(lib().fk_swiglu_backward_blocked(ptr(h13), ptr(dgt), ptr(dh13), M, HID, 128, stream()), "x"),
           bytes_=M * HID * 10, flush=False)
    # ---- GEMMs ----
    x512 = rnd(M, 512, dtype=torch.bfloat16)
    wq = rnd(1536, 512, dtype=torch.bfloat16, s=0.04)
    spec = ops.RopeSpec.from_complex(build_complex_rope_cache(32, 4096, 10000.0).to(DEV), 4096, last=True)
    report("gemm_res<ROPE> q|k|v projection + RoPE (N=1536, K=512)",
           lambda: gemm.gemm_nt(x512, wq, None, gemm.EPI_ROPE, rope=spec, rope_cols=1024, rope_S=4096), flops=2.0 * M * 1536 * 512, flush=False)
    w13 = rnd(4096, 512, dtype=torch.bfloat16, s=0.04)
    report("gemm_res<SWIGLU> w1|w3 projection + gate (N=4096, K=512)", lambda: gemm.gemm_nt(x512, w13, None, gemm.EPI_SWIGLU),
           flops=2.0 * M * 4096 * 512, flush=False, note="HBM floor of its three outputs + input: 7.0 GB = 1.07 ms")
    wp = rnd(512, 512, dtype=torch.bfloat16, s=0.04)
    report("gemm_res<STORE> out-projection (N=512, K=512)", lambda: gemm.gemm_nt(x512, wp), flops=2.0 * M * 512 * 512, flush=False)
    gated = rnd(M, HID, dtype=torch.bfloat16)
    w2 = rnd(512, HID, dtype=torch.bfloat16, s=0.02)
    report("gemm_stream w2 (N=512, K=2048)", lambda: gemm.gemm_nt(gated, w2), flops=2.0 * M * 512 * HID, flush=False)
    w13t = rnd(512, 4096, dtype=torch.bfloat16, s=0.02)
    report("gemm_stream w1|w3 input gradient (N=512, K=4096)", lambda: gemm.gemm_nt(h13, w13t), flops=2.0 * M * 512 * 4096, flush=False)
    report("gemm_tn w1|w3 weight gradient (4096 x 512 over 524288 tokens)", lambda: gemm.gemm_tn(h13, x512), flops=2.0 * M * 4096 * 512, flush=False)
    report("gemm_tn w2 weight gradient (512 x 2048)", lambda: gemm.gemm_tn(x512, gated), flops=2.0 * M * 512 * HID, flush=False)
    del h13, dgt, dh13, gated
    # ---- patch embedding ----
    sig = rnd(M // 4096, 512, 256)
    we, be = rnd(512, 32, s=0.2), rnd(512, s=0.1)
    report("gemm_tn (patch mode) patch embedding [trials, 512, 256] -> [trials x 4096, 512]", lambda: gemm.patch_embed(sig, we, be),
           bytes_=(M // 4096) * 512 * 256 * 2 + M * 512 * 2, flush=False, note="includes the fp32 -> bf16 cast of the signal (torch)")
    # ---- attention (16 trials per launch, as the step does per layer at B = 128 / 8) ----
    B, S, H = 16, 4096, 16
    qkv = rnd(B, S, 3 * H * 32, dtype=torch.bfloat16)
    mask = ops.LabelMask.block_causal(B, S, 256, DEV)
    pairs = float(mask.visible_pairs())
    wgt = rnd(B, S, H * 32, dtype=torch.bfloat16)
    from frankenstein_b200 import _lib
    for it in range(1 if ONCE else 3):
        _lib.TIMER.reset()
        _lib.TIMER.enabled = True
        xq = qkv.clone().requires_grad_(True)
        out = ops.attention_qkv(xq * 1.0, H, None, mask)
        out.backward(wgt)
        _lib.TIMER.enabled = False
        ts = _lib.TIMER.summary()
    qk = 2.0 * pairs * H * 32
    if not ONCE:
        for nm, mult, label in (("attn_fwd", 2, "attn_fwd_tc (S = Q K^T, O = P V)"), ("attn_bwd_dkv", 4, "attn_bwd_tc<DKV> (S, dP, dV, dK)"),
                                ("attn_bwd_dq", 1, "attn_bwd_tc<DQ> (dQ owed; S / dP recomputed)"), ("attn_delta", 0, "attn_aug (delta + statistics rows)")):
            n, ms, _ = ts[nm]
            rec = {"kernel": label + " 16 trials x 16 heads x 4096 tokens, block-causal", "ms": round(ms / n, 4)}
            if mult:
                rec.update(bound="tensor / MUFU", algorithmic_GF=round(mult * qk / 1e9, 1), achieved_TFps=round(mult * qk / (ms / n) / 1e9, 1),
                           frac_of_burst_peak=round(mult * qk / (ms / n) / 1e9 / TC, 3))
            else:
                by = B * S * H * (32 * 2 * 2 + 4 + 4 + 32)
                rec.update(bound="hbm", algorithmic_MB=round(by / 1e6, 1), achieved_GBps=round(by / (ms / n) / 1e6, 1),
                           frac_of_hbm_peak=round(by / (ms / n) / 1e6 / HBM, 3))
            print(json.dumps(rec), flush=True)
    del qkv, wgt
    # ---- perceiver attention ----
    q = rnd(128, 32, 64, dtype=torch.bfloat16)
    k, v = rnd(128, 4096, 64, dtype=torch.bfloat16), rnd(128, 4096, 64, dtype=torch.bfloat16)
    qa, ka, va = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = ops.small_attention(qa, ka, va, 4)
    go = rnd(128, 32, 64, dtype=torch.bfloat16)
    report("small_attn_fwd (32 queries x 4096 keys, 4 heads x 16, 128 trials)", lambda: ops.small_attention(q, k, v, 4),
           bytes_=2 * 128 * 4096 * 64 * 2, note="+ combine kernel")
    report("small_attn_bwd", lambda: torch.autograd.grad(o, (qa, ka, va), go, retain_graph=True), bytes_=4 * 128 * 4096 * 64 * 2,
           note="+ dq kernel")
    # ---- VQ ----
    N, K, Dv = 16384, 8192, 256
    e = rnd(N, Dv)
    embed = rnd(K, Dv)
    xn, xb, inv = fvq.prepare_input(e, False)
    cb, c2 = fvq.prepare_codebook(embed, False)
    report("vq_prepare_input", lambda: fvq.prepare_input(e, False), bytes_=N * Dv * 6)
    report("vq_search (tcgen05)", lambda: fvq.search(xb, cb, c2, K, False), flops=2.0 * N * K * Dv)
    cv, ci = fvq.search(xb, cb, c2, K, False)
    report("vq_finish (candidate merge, exact re-score, gather, STE, commit loss)", lambda: fvq.finish(xn, embed, cv, ci, False, True, 0.25),
           bytes_=N * Dv * 12 + cv.numel() * 8)
    ind, qz, loss = fvq.finish(xn, embed, cv, ci, False, True, 0.25)
    report("vq_ema_stats (count, scan, fill, segmented sums)", lambda: fvq.ema_stats(xn, ind, K), bytes_=N * (Dv * 4 + 8) + K * (Dv + 1) * 4)
    report("vq_perplexity", lambda: perplexity(ind, K), bytes_=N * 8 + K * 4)
    # ---- masked L1 ----
    ss = SoundStream(C=32, D=32, codebook_size=64, n_electrodes=512)
    pred = rnd(128, 512, 512, dtype=torch.bfloat16).requires_grad_(True)
    gt = rnd(128, 512, 512)
    report("masked_l1 forward", lambda: ss.custom_l1_loss(pred, gt), bytes_=128 * 512 * 512 * 6, flush=False)
    # ---- input pipeline ----
    rng = np.random.default_rng(0)
    lengths = rng.integers(400, 600, size=128)
    volt = [torch.from_numpy(rng.standard_normal((T, 256)).astype(np.float32)).to(DEV) for T in lengths]
    spk = [torch.from_numpy(rng.poisson(2.0, size=(T, 256)).astype(np.float32)).to(DEV) for T in lengths]
    blocks = rng.integers(0, 12, size=128)
    pk = dp.PackedTrials(volt, spk, blocks, DEV)
    tot_rows = int(lengths.sum())
    report("input_trial_moments x2 + block_reduce x2 (block statistics)", lambda: pk.block_stats(), bytes_=2 * tot_rows * 512 * 4)
    report("input_normalize (z-score + Gaussian + pad -> [128, 512, 512])", lambda: pk.normalize(512, True),
           bytes_=min(tot_rows, 128 * 512) * 512 * 4 + 128 * 512 * 512 * 4)
    # ---- prefix hand-off ----
    wte, wpe = rnd(50304, 768, s=0.02), rnd(1024, 768, s=0.02)
    idx = torch.randint(0, 50304, (128, 25), device=DEV)
    pre = rnd(128, 32, 768)
    report("prefix_embed_fwd (32 prefix + 25 tokens, n_embd 768, 128 trials)", lambda: prefix.prefix_embed(wte, wpe, idx, pre),
           bytes_=128 * 57 * 768 * 4 * 2 + 57 * 768 * 4, note="includes the fp32 copies of wte / wpe the wrapper makes (no-ops for fp32 parameters)")
    # ---- convolution helpers (cfg-4 SoundStream activations: 128 trials x 512 bins x 256 channels) ----
    from frankenstein_b200 import conv
    xa = rnd(128, 512, 256, dtype=torch.bfloat16)
    report("pad_rows (causal left pad k - 1 = 2, [128, 512, 256] bf16 -> padded signal)", lambda: conv._padded(xa, 514, 2, 3),
           bytes_=128 * 512 * 256 * 2 + (128 * 514 + 3) * 256 * 2)
    gy = rnd(128 * 514, 256, dtype=torch.bfloat16)
    report("colsum_partials + reduce (bias gradient of a convolution, [65792, 256] bf16)", lambda: ops.column_sum(gy),
           bytes_=128 * 514 * 256 * 2)
    gl = rnd(M, 512, dtype=torch.bfloat16)
    report("colsum_partials + reduce (bias gradient of a Linear, [M, 512] bf16)", lambda: ops.column_sum(gl), bytes_=M * 512 * 2)


if __name__ == "__main__":
    main()
