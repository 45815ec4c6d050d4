"""Time the tcgen05 search kernel (and the other VQ kernels) with CUDA events; prints TFLOP/s and GB/s."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from frankenstein_b200 import vector_quantize as fvq
from frankenstein_b200._lib import lib, ptr, stream, check


def time_fn(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    shapes = [(4096, 512, 64)] + [(16384, k, 256) for k in (512, 1024, 2048, 4096, 8192, 16384, 32768, 65536)]
    if "--shape" in sys.argv:
        i = sys.argv.index("--shape")
        shapes = [tuple(int(v) for v in sys.argv[i + 1:i + 4])]
    out = []
    for N, K, D in shapes:
        X = torch.randn(N, D, device="cuda"); C = torch.randn(K, D, device="cuda")
        xn, xb, _ = fvq.prepare_input(X, False)
        cb, c2 = fvq.prepare_codebook(C, False)
        med, best = time_fn(lambda: fvq.search(xb, cb, c2, K, False))
        fl = 2.0 * N * K * D
        cv, ci = fvq.search(xb, cb, c2, K, False)
        m2, b2 = time_fn(lambda: fvq.finish(xn, C, cv, ci, False, True, 0.25))
        ind, q, l = fvq.finish(xn, C, cv, ci, False, True, 0.25)
        m3, b3 = time_fn(lambda: fvq.ema_stats(xn, ind, K))
        rec = dict(N=N, K=K, D=D, search_ms_med=med, search_ms_best=best, tflops_med=fl / med / 1e9, tflops_best=fl / best / 1e9,
                   finish_ms=m2, ema_stats_ms=m3, S=cv.shape[1])
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/time_search.json", "w"), indent=1)


if __name__ == "__main__":
    main()
