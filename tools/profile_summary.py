#!/usr/bin/env python
"""Turns the scratch ncu reports of one GPU round (gpurun_out/<tag>/) into the committed summaries under profiles/:

    python tools/profile_summary.py <tag> <round-label> [out-dir]

  profiles/<label>_launches.csv          ncu launch list of `bench.py --steps 2 --warmup 1`
  profiles/<label>_ncu_full_summary.md   key counters of the `ncu --set full` captures (dense + block-diagonal kernels)
  profiles/<label>_dense_by_line.txt     per-source-line instruction / stall-sample / shared-memory breakdown
  profiles/<label>_bench.json            the bench line of the same round
  profiles/traffic.json                  dram bytes per launch of the bench-sized kernels (read by bench.py)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__cycles_elapsed.max",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    tag, label = sys.argv[1], sys.argv[2]
    src = os.path.join(ROOT, "gpurun_out", tag)
    # run ON the GPU box (the .ncu-rep files are too large to travel back): third argument = output directory
    dst = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    md = [f"# {label}: `ncu --set full --clock-control none` counters (B200, sm_100a)", "",
          "Raw reports stay in gpurun_out/ (scratch).  Per-launch times under ncu are serialised and cold-cache: "
          "compare shares, not absolutes; the bench line's own CUDA-event timing is in the *_bench.json beside this file.", ""]
    traffic = {}
    for rep, steps_note in (("dense_full", "dense"), ("bd_full", "blockdiag"), ("dense_full_100k", "dense, bench-sized launch"),
                            ("twist_full", "camera-twist pipeline, 20 k trials x 440 frames")):
        path = os.path.join(src, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        hdr, units, vals = raw(path)
        for v in vals:
            name = v[hdr.index("Kernel Name")]
            grid = v[hdr.index("launch__grid_size")]
            md += [f"## {name}  ({steps_note}; capture {rep})", "", "| metric | value | unit |", "|---|---|---|"]
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    md.append(f"| {w} | {v[i]} | {units[i]} |")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v[i] or 0) > 0.05:
                    md.append(f"| {h} | {v[i]} | cycles per issued instruction |")
            md.append("")
            if rep == "dense_full":
                # hardware-counted FP64 flops per filter-step of the 14 800-trial x 300-frame capture (must sit
                # within [0.5, 1.1] x F_alg): 2 dfma + dmul + dadd thread instructions, per elapsed cycle x elapsed cycles
                try:
                    per_cycle = sum(k * float(v[hdr.index(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")])
                                    for op, k in (("dfma", 2.0), ("dmul", 1.0), ("dadd", 1.0)))
                    cycles = float(v[hdr.index("smsp__cycles_elapsed.max")]) if "smsp__cycles_elapsed.max" in hdr \
                        else float(v[hdr.index("sm__cycles_elapsed.max")])
                    traffic["dense_executed_flops_per_filter_step"] = per_cycle * cycles / (14800 * 300)
                    md.insert(len(md) - 1, f"| executed FP64 flops per filter-step (2 dfma + dmul + dadd) | "
                                           f"{traffic['dense_executed_flops_per_filter_step']:.0f} | flop |")
                except (ValueError, KeyError):
                    pass
            if rep == "bd_full":
                # the block-diagonal family's OWN executed flop count (never F_alg): bench-sized launch, 100 000 x 300 steps
                try:
                    per_cycle = sum(k * float(v[hdr.index(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")])
                                    for op, k in (("dfma", 2.0), ("dmul", 1.0), ("dadd", 1.0)))
                    cycles = float(v[hdr.index("smsp__cycles_elapsed.max")]) if "smsp__cycles_elapsed.max" in hdr \
                        else float(v[hdr.index("sm__cycles_elapsed.max")])
                    traffic["blockdiag_executed_flops_per_filter_step"] = per_cycle * cycles / (100000 * 300)
                    md.insert(len(md) - 1, f"| executed FP64 flops per filter-step (2 dfma + dmul + dadd) | "
                                           f"{traffic['blockdiag_executed_flops_per_filter_step']:.0f} | flop |")
                except (ValueError, KeyError):
                    pass
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            if rep not in ("dense_full", "twist_full"):            # bench-sized launches only
                key = "dense_quad_kernel<8, GMCKF>" if "dense_quad" in name else \
                    ("dense" if "dense" in name else "blockdiag") + "_closed_loop_kernel<8, GMCKF>"
                traffic[key] = to_bytes(v[ir], units[ir]) + to_bytes(v[iw], units[iw])
    with open(os.path.join(dst, f"{label}_ncu_full_summary.md"), "w") as fh:
        fh.write("\n".join(md) + "\n")
    for a, b in (("launches.csv", f"{label}_launches.csv"), ("bench.json", f"{label}_bench.json")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(dst, b))
    if traffic:
        with open(os.path.join(dst, "traffic.json"), "w") as fh:
            json.dump(traffic, fh, indent=1)
    rep = os.path.join(src, "dense_full.ncu-rep")
    if os.path.exists(rep):
        sass = os.path.join("/tmp", f"{label}_sass.csv")
        with open(sass, "w") as fh:
            subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=fh, stderr=subprocess.DEVNULL)
        # per-phase view of the quad kernel (regions between its named barriers = the hand-over points of the two roles)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_regions.py"), sass, str(14800 * 300)],
                             capture_output=True, text=True).stdout
        with open(os.path.join(dst, f"{label}_dense_by_region.txt"), "w") as fh:
            fh.write(out)
        cub = "/tmp/uvs_cubin"
        os.makedirs(cub, exist_ok=True)
        subprocess.run(["cuobjdump", "-xelf", "dense_quad_m8", os.path.join(ROOT, "uncalibrated-visual-servoing_b200", "libuvs_b200.so")],
                       cwd=cub, capture_output=True)
        cubin = [f for f in os.listdir(cub) if f.startswith("dense_quad_m8")]
        if cubin:
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), sass,
                                  os.path.join(cub, cubin[0]), "dense_quad_kernelILi8ELi5", "40"],
                                 capture_output=True, text=True).stdout
            with open(os.path.join(dst, f"{label}_dense_by_line.txt"), "w") as fh:
                fh.write(out)
    print("wrote profiles/", label)


if __name__ == "__main__":
    main()
