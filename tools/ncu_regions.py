#!/usr/bin/env python
"""Per-phase view of an `ncu --set full --import-source on` capture of the quad kernel (csrc/dense_quad.cuh):

    ncu -i dense_full.ncu-rep --page source --csv > source.csv
    python tools/ncu_regions.py source.csv <filter-steps of the launch>

The kernel's instruction stream is cut at its named barriers / function returns, which are exactly the hand-over points of
the two warp roles; for every region: executed warp-instructions, FP64 instructions and shared-memory wavefronts per
filter-step, its share of the stall samples and the three largest stall reasons."""
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    steps = float(sys.argv[2])
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0

    stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    tot_s = sum(f(r, "# Samples") for r in data) or 1.0
    tot_w = sum(f(r, "L1 Wavefronts Shared") for r in data)
    tot_i = sum(f(r, "Instructions Executed") for r in data)
    print(f"{tot_i / steps:.0f} warp-instructions and {tot_w / steps:.0f} shared-memory wavefronts per filter-step "
          f"({sum(f(r, 'L1 Wavefronts Shared Ideal') for r in data) / steps:.0f} ideal)")
    print("region (first instruction ... closing barrier) | inst/step | FP64/step | smem wavefronts/step (ideal) | samples | top stalls")
    cur = []
    for r in data:
        cur.append(r)
        src = r[ix["Source"]]
        if re.search(r"\bBAR\.|\bRET\b|\bEXIT\b", src):
            inst = sum(f(x, "Instructions Executed") for x in cur)
            if inst / steps >= 2.0:
                s = sum(f(x, "# Samples") for x in cur)
                fp = sum(f(x, "Instructions Executed") for x in cur if re.search(r"\bD(FMA|MUL|ADD)\b", x[ix["Source"]]))
                w = sum(f(x, "L1 Wavefronts Shared") for x in cur)
                wi = sum(f(x, "L1 Wavefronts Shared Ideal") for x in cur)
                st = sorted(((k[6:], sum(f(x, k) for x in cur)) for k in stalls), key=lambda kv: -kv[1])[:3]
                end = re.sub(r"\s+", " ", src.strip())[:34]
                print(f"{len(cur):5d} instr ... {end:34s} | {inst / steps:7.1f} | {fp / steps:6.1f} | {w / steps:6.1f} ({wi / steps:6.1f}) | "
                      f"{100 * s / tot_s:5.1f} % | " + ", ".join(f"{k} {100 * v / max(s, 1):.0f} %" for k, v in st))
            cur = []


if __name__ == "__main__":
    main()
