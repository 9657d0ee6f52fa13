"""Per-phase cycle breakdown of the dense closed-loop frame loop (profiling build of the library only).

    UVS_PHASE_CLOCKS=1 python uncalibrated-visual-servoing_b200/build.py --out libuvs_b200_clk.so
    UVS_B200_LIB=$PWD/uncalibrated-visual-servoing_b200/libuvs_b200_clk.so python tools/phase_clocks.py

Every warp accumulates clock64() deltas between the markers UVS_CLK(i) of csrc/dense_kernels.cuh / dense.cuh; the totals
are per-warp wall cycles INCLUDING the time the warp waited for its turn on the scheduler (2 warps per scheduler), so
their sum is ~2x the per-scheduler cycles per filter-step.  The clock reads themselves perturb the schedule a little.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uvs_b200 as uvs  # noqa: E402

NAMES = ["0 plant: sincos, point in camera, pixel", "1 noise (Philox alpha-stable, 4 frames at once) + hold", "2 predict P+=Q, A = P H^T",
         "3 innovation, kernel weights (exp), 1/(c+1e-6)", "4 S build + 8x8 Cholesky (redundant per lane)", "5 gain rows K = A S^-1, X += K nu, U",
         "6 IBVS stage 1: Gram J^T J, J^T b (+ rank update k 0..1)", "7 IBVS stage 2: 6x6 Cholesky, solve, residual (+ k 2..5)",
         "8 IBVS stage 3: J^T r (+ k 6)", "9 IBVS stage 4: refinement solve (+ k 7, diagonal blocks)", "10 (fallback check)",
         "11 frame tail: dq, finiteness, integrals, integrate"]


QUAD_CHAIN = ["hold logic, innovation, kernel weights, b", "wait (1): A, S0", "S, L D L^T, S^-1, y", "wait (2)", "gain rows K, U",
              "wait (3): X, Gram matrix", "pseudo-inverse", "command, finiteness, publish", "wait (4)", "log, integrals, integrate; plant of the next frame"]
QUAD_COV = ["predict, A, S0", "wait (1)", "wait (2): S^-1, y", "X += A y, Gram, noise batch (every 4th frame), wait (3)", "rank update", "wait (4): pinv"]


def main():
    L, E = uvs._lib, uvs.engine
    lib = L.load()
    fn = lib.uvs_debug_phase_clocks
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    B, T = 29600, 300
    sweep = [float(v) for v in np.linspace(1, 2, 12)]
    p = E.filter_params(L.METHOD_GMCKF, 4, mode=L.MODE_DENSE, kernel_bw=10.0, annealing=True)
    cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=123456, alpha=1.5), T=T, dt=0.05, ibvs_gain=0.2,
                            q_start=[0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0],
                            desired_f=[149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0], sweep=sweep,
                            sweep_epoch=B // 12 + 1, sweep_param=2)
    out = E.closed_loop(cfg, 0, B)
    buf = (C.c_ulonglong * 40)()
    fn(buf, 1)
    out = E.closed_loop(cfg, 0, B, out=out)
    fn(buf, 1)
    steps = int(out["frames"].sum().item())
    allc = np.array(list(buf), dtype=float)
    tot = allc[:12] / steps
    if tot.sum() > 0:
        print(f"{steps} filter-steps; per-warp cycles per filter-step: {tot.sum():.0f}")
        for n, v in zip(NAMES, tot):
            print(f"  {v:8.0f}  {100 * v / tot.sum():5.1f} %  {n}")
    if allc[12:].sum() > 0:          # the quad kernel (dense_quad.cuh): cycles per FRAME of a four-filter CTA, per role
        frames = steps / 4.0
        ch, cv = allc[12:22] / frames, allc[22:28] / (4.0 * frames)
        print(f"quad kernel: {frames:.0f} CTA-frames; chain warp {ch.sum():.0f} cycles per frame, covariance warp {cv.sum():.0f}")
        for n, v in zip(QUAD_CHAIN, ch):
            print(f"  chain  {v:8.0f}  {100 * v / ch.sum():5.1f} %  {n}")
        for n, v in zip(QUAD_COV, cv):
            print(f"  covar  {v:8.0f}  {100 * v / cv.sum():5.1f} %  {n}")


if __name__ == "__main__":
    main()
