#!/usr/bin/env python
"""Closed-loop throughput of every estimator in both kernel families (needs a GPU):
    python tools/method_table.py [--trials 50000]
alpha-stable sweep of bench.py (BASELINE configs[2] workload), 4 / 3 / 1 features, T = 300."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import uvs_b200 as uvs

L, E = uvs._lib, uvs.engine
DESIRED = {4: bench.DESIRED_F, 3: [153.5, 144.5, 125.5, 116.5, 99.5, 144.5], 1: [128.0, 128.0]}


def rate(method, nf, mode, B):
    p = E.filter_params(method, nf, mode=mode, kernel_bw=10.0, annealing=(method != L.METHOD_KF))
    cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=bench.NOISE_SEED, alpha=1.5), T=300, dt=0.05,
                            ibvs_gain=0.2, q_start=bench.Q_START, desired_f=DESIRED[nf], pose_seed=bench.POSE_SEED,
                            sweep=bench.SWEEP, sweep_epoch=B // 12 + 1, sweep_param=2)
    out = E.closed_loop(cfg, 0, B)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    E.closed_loop(cfg, 0, B, out=out)
    e1.record()
    torch.cuda.synchronize()
    return int(out["frames"].sum().item()) / (e0.elapsed_time(e1) * 1e-3), float((out["status"] != 0).float().mean().item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=50000)
    a = ap.parse_args()
    names = {L.METHOD_KF: "KF", L.METHOD_MCKF: "MCKF", L.METHOD_IMCCKF: "IMCC-KF", L.METHOD_GMCKF: "RMCKF"}
    print("| features | estimator | dense (general P, Q = R = I) steps/s | block-diagonal steps/s | failed trials |")
    print("|---|---|---|---|---|")
    for nf in (4, 3, 1):
        for m, nm in names.items():
            d, fd = rate(m, nf, L.MODE_DENSE, a.trials if m != L.METHOD_MCKF else a.trials // 4)
            b, fb = rate(m, nf, L.MODE_BLOCKDIAG, a.trials)
            print(f"| {nf} | {nm} | {d:.3g} | {b:.3g} | {fd:.4f} / {fb:.4f} |")


if __name__ == "__main__":
    main()
