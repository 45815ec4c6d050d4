"""Per-kernel share of one bench step from an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python scripts/launch_summary.py launches.csv [n_last_steps_marker]  -- takes the launches after the
LAST occurrence of the step's first kernel pattern (the optimizer zero_grad is not a kernel, so we split on the
VQ-VAE encoder's first conv instead: the last `vq_prepare_input` launch marks the last step)."""
import collections
import csv
import sys


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        us = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        rows.append((r[ik], us))
    # last step = from the last forward's first fk kernel backwards to the previous optimizer: approximate by taking
    # everything after the second-to-last vq_prepare_input launch up to the last one's successor optimizer kernels.
    idx = [i for i, (k, _) in enumerate(rows) if "vq_prepare_input" in k]
    # two timed regions (resident + e2e) each with 1 step follow 3 warm-ups: use the step that starts at idx[-2]
    start = idx[-2] if len(idx) >= 2 else 0
    # the step begins with the encoder convs that precede vq_prepare_input: back up to the preceding AdamW kernel
    s0 = start
    while s0 > 0 and "adam" not in rows[s0 - 1][0].lower():
        s0 -= 1
    step = rows[s0:idx[-1]]
    e0 = len(step)
    # cut at the end of that step's optimizer
    last_adam = max((i for i, (k, _) in enumerate(step) if "adam" in k.lower()), default=e0 - 1)
    step = step[:last_adam + 1]
    tot = sum(us for _, us in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, us in step:
        name = k.split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += us
    print(f"launches in step: {len(step)}   sum of kernel durations: {tot / 1e3:.2f} ms (cold-cache, serialised under ncu)")
    print(f"{'kernel':72s} {'n':>5s} {'ms':>9s} {'share':>7s}")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{name:72s} {n:5d} {us / 1e3:9.3f} {100 * us / tot:6.2f}%")
    ours = sum(us for k, us in step if k.startswith("fk::") or "fk::" in k or k.startswith("vq_") or k.startswith("attn_"))
    print(f"share of frankenstein_b200 kernels (fk::*): {100 * ours / tot:.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
