#!/usr/bin/env python
"""Benchmark of the neural-encoder training step (BASELINE.json metric: trials/sec of the VQ + Brainformer
train step; VQ-search fraction of the bf16 tensor-core peak).

    python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle port)

Workload "cfg4-joint" (per GPU): 128 synthetic trials x[512 bins, 512 ch] ->
SoundStream(C=256, D=256, K=8192, Euclidean) on all 512 channels  +  BrainFormer (Encoder window 512, 256
electrodes, patch 32 -> 4096 tokens/trial, dim 512, 4 layers, 16 heads x 32, hidden 2048, + perceiver) on the
first 256 channels; loss = sum of both; one backward; value-clip(1.0); fused AdamW.  Weak scaling: 128 trials
per GPU.  One step = one pass over one batch.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist
import torch.nn as nn

TRIALS_PER_GPU = 128
T_BINS, N_CH = 512, 512
VQ_K, VQ_D, VQ_C = 8192, 256, 256
CPU_SAMPLE_TRIALS = 2
WORKLOAD = ("cfg4-joint: SoundStream(C=256,D=256,K=8192,euclid) on 512 ch + BrainFormer(Encoder window 512, "
            "256 electrodes, patch 32 -> 4096 tokens/trial, dim 512, 4 layers, 16x32 heads, hidden 2048; "
            "perceiver 32 tokens) on 256 ch; AdamW + value clip")
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
# (attention rows: the capture in profiles/r01_attention_tc_ncu_full.txt is a 16-trial launch; trials are independent,
#  so a B-trial launch moves B/16 times those bytes)
NCU_TRAFFIC = {"vq_search": 12660224}
NCU_TRAFFIC_PER_16_TRIALS = {"attn_fwd": 202000896 + 54899712, "attn_bwd_dkv": 277223424 + 112426240,
                             "attn_bwd_dq": 277200640 + 58217984}


def model_configs():
    from frankenstein_b200.brainformer import Config, MAEConfig
    enc = MAEConfig(window_size=512, n_electrodes=256, patch_size=32, dim=512, n_layers=4, head_dim=32, hidden_dim=2048,
                    n_heads=16, n_kv_heads=16, n_dec_layers=4, decoder_dim=512)
    per = Config(encoder=enc, n_output_tokens=32, output_dim=768, dim=512, n_layers=2, head_dim=16, hidden_dim=512,
                 n_heads=4, n_kv_heads=4)
    return enc, per


class Joint(nn.Module):
    """loss = SoundStream(x)[0] + BrainFormer(x[..., :256], targets)[0]   (SURVEY.md section 8 'joint' definition)."""

    def __init__(self):
        super().__init__()
        from frankenstein_b200.brainformer import BrainFormer
        from frankenstein_b200.vq_brain import SoundStream
        _, per = model_configs()
        self.vqvae = SoundStream(C=VQ_C, D=VQ_D, codebook_size=VQ_K, n_electrodes=N_CH, use_cosine_sim=False)
        self.brainformer = BrainFormer(per)

    def forward(self, x, targets):
        l_vq, _ = self.vqvae(x)
        l_bf, _ = self.brainformer(x[..., :256].contiguous(), targets)
        return l_vq + l_bf


def synth_batch(n, seed):
    """SURVEY 8d synthetic trials: 5-tap smoothed noise, per-channel z-score, zero-padded tail."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, T_BINS + 4, N_CH, generator=g)
    x = (x[:, 0:T_BINS] + x[:, 1:T_BINS + 1] + x[:, 2:T_BINS + 2] + x[:, 3:T_BINS + 3] + x[:, 4:T_BINS + 4]) / 5.0
    x = (x - x.mean(dim=1, keepdim=True)) / x.std(dim=1, keepdim=True)
    pad = torch.randint(0, T_BINS // 4 + 1, (n,), generator=g)
    for b in range(n):
        if pad[b] > 0:
            x[b, T_BINS - int(pad[b]):] = 0
    t = torch.randn(n, 32, 768, generator=g)
    return x.contiguous(), t.contiguous()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 6 or not parts[0].isdigit():
                    continue
                sm.append(int(parts[0]))
                mx.append(int(parts[1]))
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons)}
        return out


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own CPU path restated in oracle/ (the reference tree is not on the GPU box)
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, trials=CPU_SAMPLE_TRIALS):
    from frankenstein_b200.brainformer import BrainFormer
    from frankenstein_b200.vq_brain import SoundStream
    from oracle import brainformer_ref
    from oracle.soundstream_ref import SoundStreamRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    enc, per = model_configs()
    ss_sd = SoundStream(C=VQ_C, D=VQ_D, codebook_size=VQ_K, n_electrodes=N_CH, use_cosine_sim=False).state_dict()
    bf_sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "attn_mask" not in k)
             for k, v in BrainFormer(per).state_dict().items()}
    ss = SoundStreamRef(ss_sd, VQ_D, VQ_K, use_cosine_sim=False, training=True)
    enc_cfg = dict(window_size=512, n_electrodes=256, patch_size=32, head_dim=32, n_heads=16)
    per_cfg = dict(head_dim=16, n_output_tokens=32, n_heads=4)
    params = ss.parameters() + [v for v in bf_sd.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01)
    x, t = synth_batch(trials, 99)

    def step():
        opt.zero_grad(set_to_none=True)
        l_vq, _ = ss(x)
        l_bf, _ = brainformer_ref.brainformer_forward(bf_sd, x[..., :256].contiguous(), enc_cfg, per_cfg, t)
        loss = l_vq + l_bf
        loss.backward()
        torch.nn.utils.clip_grad_value_(params, 1.0)
        opt.step()
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": trials * steps / dt, "unit": "trials/s", "cores": cores, "kind": "port",
            "sample": f"{steps} timed steps of {trials} trial(s) after {warmup} warm-up, fp32, torch {torch.__version__} "
                      f"on {cores} host threads, oracle/ restatement of models/vq_brain.py + models/brainformer.py"}, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    base, ms = cpu_reference_run(steps, warmup)
    line = {"impl": "reference", "metric": "trials/sec VQ+Brainformer train step", "value": base["value"], "unit": "trials/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "trials_per_gpu": TRIALS_PER_GPU, "bins": T_BINS,
                       "sample": f"bounded CPU sample of the same step: {CPU_SAMPLE_TRIALS} trials per timed step, scaled to trials/s"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--trials", type=int, default=TRIALS_PER_GPU, help="trials per GPU (the metric's config is 128)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from frankenstein_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the sm_100a kernels have no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = args.steps
    B = args.trials

    torch.manual_seed(1234)       # identical initial weights on every rank
    model = Joint().to(dev).train()
    net = model
    if world > 1:
        # VQ buffers are kept identical by the packed EMA all-reduce inside the quantiser, not by DDP broadcasts
        net = nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], broadcast_buffers=False,
                                                  gradient_as_bucket_view=True)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True)

    # host-side pinned batches (distinct per rank) and a device-resident pool (each batch 134 MB > the 126 MB L2)
    n_pool = 3
    host = [synth_batch(B, 1234 + rank * 100 + i) for i in range(n_pool)]
    host = [(x.pin_memory(), t.pin_memory()) for x, t in host]
    pool = [(x.to(dev), t.to(dev)) for x, t in host]
    h2d_bytes = host[0][0].numel() * 4 + host[0][1].numel() * 4

    def step(x, t):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = net(x, t)
        loss.backward()
        torch.nn.utils.clip_grad_value_(params, 1.0)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(W):
        loss = step(*pool[i % n_pool])
    barrier()
    assert torch.isfinite(loss).all(), "non-finite loss in warm-up"

    # ---- timed region 1: inputs resident in HBM ----
    clocks = ClockSampler(local_rank)
    clocks.start()
    _lib.reset_launch_count()
    _lib.TIMER.reset()
    _lib.TIMER.enabled = True
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = step(*pool[i % n_pool])
    e1.record()
    barrier()
    _lib.TIMER.enabled = False
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count()
    ksum = _lib.TIMER.summary()
    clk = clocks.stop()
    final_loss = float(loss.detach())

    # ---- timed region 2: end to end (pinned host batch -> device every step, loss read back every step) ----
    xd = torch.empty_like(pool[0][0])
    td = torch.empty_like(pool[0][1])
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(K):
        hx, ht = host[i % n_pool]
        xd.copy_(hx, non_blocking=True)
        td.copy_(ht, non_blocking=True)
        loss = step(xd, td)
        _ = loss.item()                          # the trainer logs the loss every step (utils/train_utils.py:147)
    t1.record()
    barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tc_peak = peaks.get("bf16_tflops_sustained", 1400.0)       # kernel timed inside a long step -> sustained figure
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        n_vq = B * (T_BINS // 4)
        # ---- per-kernel rooflines from CUDA events recorded around each launch inside the timed region ----
        dens = (16 + 1) / (2.0 * 16)               # block-causal density with 16 time patches
        S, H, HD = 4096, 16, 32
        qk = 2.0 * B * H * S * S * HD * dens       # flops of ONE score-shaped matmul over the visible pairs
        # algorithmic flops per launch: fwd = QK^T + PV; dK/dV kernel = S, dP, dV, dK; the dQ kernel owes only dQ
        # (its S / dP re-computation is overhead of the two-kernel split and is not counted as useful work)
        alg = {"vq_search": 2.0 * n_vq * VQ_K * VQ_D, "attn_fwd": 2 * qk, "attn_bwd_dkv": 4 * qk, "attn_bwd_dq": 1 * qk}
        names = {"vq_search": "fk::vq_search_kernel (tcgen05/TMEM/TMA nearest-codeword search)",
                 "attn_fwd": "fk::attn_fwd_tc_kernel (label-mask flash attention forward, tcgen05/TMEM/TMA)",
                 "attn_bwd_dkv": "fk::attn_bwd_tc_kernel<DKV> (label-mask flash attention dK/dV, tcgen05/TMEM/TMA)",
                 "attn_bwd_dq": "fk::attn_bwd_tc_kernel<DQ> (label-mask flash attention dQ, tcgen05/TMEM/TMA)"}
        # ncu --set full captures (profiles/): dram__bytes_read.sum + dram__bytes_write.sum per launch
        traffic = {"vq_search": NCU_TRAFFIC.get("vq_search")}
        traffic.update({k: int(v * B / 16) for k, v in NCU_TRAFFIC_PER_16_TRIALS.items()})
        kern = {}
        for name, flops in alg.items():
            if name not in ksum:
                continue
            n, ms, _ = ksum[name]
            big = [1] if name == "vq_search" else None
            avg_ms = ms / n
            ach = flops / (avg_ms * 1e-3) / 1e12
            kern[name] = {"kernel": names[name], "bound": "tensor", "achieved": ach, "peak": tc_peak, "unit": "TFLOP/s",
                          "frac": ach / tc_peak, "traffic": traffic.get(name), "algorithmic_flops_per_launch": flops,
                          "avg_launch_ms": avg_ms, "launches_timed": n, "ms_per_step": ms / K, "peak_source": peak_src}
        # the roofline object is the kernel with the largest share of the step
        roof = None
        if kern:
            dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
            roof = dict(kern[dom])
            roof["share_of_step"] = roof["ms_per_step"] / (ms_total / K)
        total_trials = B * world * K
        line = {
            "metric": "trials/sec VQ+Brainformer train step", "value": total_trials / (ms_total * 1e-3), "unit": "trials/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "trials_per_gpu": B, "global_batch": B * world, "bins": T_BINS, "parallelism": f"dp{world}",
                       "l2_policy": "inputs larger than L2 (134 MB batch rotated over a 3-batch pool)"},
            "clocks": clk,
            "e2e": {"value": total_trials / (ms_e2e * 1e-3), "unit": "trials/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches),
            "loss": final_loss,
            "roofline": roof,
            "vq_search": kern.get("vq_search"),
            "kernels": kern,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_reference_run(steps=2, warmup=1)
            except Exception as e:  # the number is informational; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": "trials/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
