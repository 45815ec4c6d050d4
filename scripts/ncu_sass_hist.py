"""Opcode histogram (executed warp-instructions and stall samples) from `ncu --page source --csv`."""
import collections
import csv
import subprocess
import sys


def main(path, top=30):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    iT = hdr.index("Avg. Threads Executed")
    h, hs = collections.Counter(), collections.Counter()
    data = []
    for r in rows[2:]:
        if len(r) <= iW or not r[iE].isdigit():
            continue
        e, w, s = int(r[iE]), int(r[iW] or 0), r[iS].strip()
        toks = s.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        h[op] += e
        hs[op] += w
        data.append((e, w, s, r[iT]))
    tot, tots = sum(h.values()), sum(hs.values())
    print(f"total warp-instructions {tot}, stall samples {tots}")
    for op, c in h.most_common(top):
        print(f"{op:30s} {c:11d} {100 * c / tot:5.1f}%   samples {hs[op]:7d} {100 * hs[op] / max(tots, 1):5.1f}%")
    print("---- top stall lines")
    for e, w, s, t in sorted(data, key=lambda d: -d[1])[:25]:
        print(f"{w:7d} {e:10d} thr={t:>5s}  {s[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
