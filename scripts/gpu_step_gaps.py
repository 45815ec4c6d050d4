"""GPU timeline of one training step (torch.profiler / CUPTI): busy time, idle gaps and the kernels that precede the gaps.
Diagnostic only -- a number taken under a profiler is never a bench value.  usage: python scripts/gpu_step_gaps.py [workload]"""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4-joint"
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = bench.build_model(wl).to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True)
B = bench.DEFAULT_TRIALS[wl]
pool = [tuple(a.to(dev) for a in bench.synth_batch(B, 1234 + i)) for i in range(2)]

def step(x, t):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model(x, t)
    loss.backward()
    torch.nn.utils.clip_grad_value_(params, 1.0)
    opt.step()
    return loss

for i in range(4):
    step(*pool[i % 2])
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        step(*pool[i % 2])
    torch.cuda.synchronize()
path = "/tmp/fk_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
# the middle step: between the 1st and 2nd fused-AdamW launch groups
adam = [i for i, e in enumerate(ev) if "multi_tensor_apply" in e["name"] and "Adam" in e["name"]]
cut = [i for k, i in enumerate(adam) if k == 0 or i - adam[k - 1] > 50]      # first adam kernel of each step
lo, hi = cut[0], cut[1]
evs = ev[lo:hi]
t0, t1 = evs[0]["ts"], evs[-1]["ts"] + evs[-1]["dur"]
busy, end, gaps = 0.0, evs[0]["ts"], []
for i, e in enumerate(evs):
    s, d = e["ts"], e["dur"]
    if s > end:
        gaps.append((s - end, evs[i - 1]["name"][:70] if i else "-", e["name"][:70]))
        busy += d
    else:
        busy += max(0.0, s + d - end)
    end = max(end, s + d)
print(f"workload {wl}: launches {len(evs)}  span {(t1 - t0) / 1e3:.2f} ms  busy {busy / 1e3:.2f} ms  idle {(t1 - t0 - busy) / 1e3:.2f} ms  gaps {len(gaps)}")
h = collections.Counter()
for g, a, b in gaps:
    h["<2us" if g < 2 else "2-5us" if g < 5 else "5-20us" if g < 20 else "20-100us" if g < 100 else ">100us"] += g
print("idle by gap size (us):", {k: round(v, 1) for k, v in h.items()})
by = collections.defaultdict(float)
for g, a, b in gaps:
    by[(a, b)] += g
for (a, b), g in sorted(by.items(), key=lambda kv: -kv[1])[:25]:
    print(f"{g:9.1f} us  after {a}  ->  {b}")
