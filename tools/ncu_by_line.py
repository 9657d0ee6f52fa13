#!/usr/bin/env python
"""Aggregates an ncu SASS source page per CUDA source line / inlined function.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_by_line.py sass.csv <cubin> <mangled kernel substring> [top]

Joins ncu's per-SASS-instruction counters with `nvdisasm -g` line info (needs -lineinfo at
compile time) and prints, per source line, executed warp instructions and stall samples.
"""
import collections
import csv
import re
import subprocess
import sys


def line_map(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    amap, cur, infn, stack = {}, None, False, None
    for ln in out.splitlines():
        if ln.startswith(".text."):
            infn = kernel in ln
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (m.group(1).split("/")[-1], int(m.group(2)), tuple((a.split("/")[-1], int(b)) for a, b in inl))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            amap[int(m.group(1), 16)] = (cur, m.group(2))
    return amap


def main():
    sass_csv, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    amap = line_map(cubin, kernel)
    rows = list(csv.reader(open(sass_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed",
                                    "Thread Instructions Executed", "L1 Wavefronts Shared",
                                    "L1 Wavefronts Shared Ideal", "L2 Theoretical Sectors Local")}
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    per_outer = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    tot = [0, 0, 0, 0, 0, 0]
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None:
            base = addr
        key = amap.get(addr - base)
        vals = [int(float(r[ci[n]] or 0)) for n in ("Instructions Executed", "# Samples", "Thread Instructions Executed",
                                                   "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal",
                                                   "L2 Theoretical Sectors Local")]
        loc = key[0] if key else ("?", 0, ())
        outer = loc[2][-1] if loc[2] else (loc[0], loc[1])
        for acc in (per_line[(loc[0], loc[1])], per_outer[outer], tot):
            for i, v in enumerate(vals):
                acc[i] += v
    def show(title, d):
        print(f"\n== {title}  (inst%, samples%, avg threads, smem wavefronts/ideal, local sectors)")
        for k, v in sorted(d.items(), key=lambda kv: -kv[1][1])[:top]:
            print(f"{k[0]:>24}:{k[1]:<5} inst {100*v[0]/max(tot[0],1):5.1f}%  samp {100*v[1]/max(tot[1],1):5.1f}%  "
                  f"thr {v[2]/max(v[0],1):4.1f}  smem {v[3]:>11}/{v[4]:<11} local {v[5]}")
    print("total inst", tot[0], "samples", tot[1])
    show("by line in the kernel body (outermost inlining site)", per_outer)
    show("by innermost source line", per_line)


if __name__ == "__main__":
    main()
