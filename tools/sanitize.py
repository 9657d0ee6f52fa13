#!/usr/bin/env python
"""Small launches of every kernel family, to run under compute-sanitizer on a GPU box:

    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py        (shared-memory hazards between lanes)
    compute-sanitizer --tool synccheck python tools/sanitize.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np          # noqa: E402
import torch                # noqa: E402
import uvs_b200 as uvs      # noqa: E402

L, E = uvs._lib, uvs.engine
Q_START = [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0]
DESIRED = {4: [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0], 3: [153.5, 144.5, 125.5, 116.5, 99.5, 144.5],
           1: [128.0, 128.0]}


def spd(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n))
    return np.eye(n) + 0.3 * (a @ a.T) / n


def main():
    B, T = 24, 12
    noise = E.noise_params(L.NOISE_ALPHA_STABLE, seed=1, alpha=1.4)
    for nf in (4, 3, 1):
        m, N = 2 * nf, 12 * nf
        for method in (L.METHOD_KF, L.METHOD_MCKF, L.METHOD_IMCCKF, L.METHOD_GMCKF):
            for mode, kw in ((L.MODE_DENSE, {}), (L.MODE_BLOCKDIAG, {}),
                             (L.MODE_DENSE, dict(Q=spd(N, 1), R=spd(m, 2)))):
                p = E.filter_params(method, nf, mode=mode, annealing=True, **kw)
                cfg = E.closed_loop_cfg(p, noise, T=T, dt=0.05, ibvs_gain=0.2, q_start=Q_START, desired_f=DESIRED[nf],
                                        P0=spd(N, 3) if kw else None, servo_gain=0.65 if kw else 0.0)
                E.closed_loop(cfg, 0, B, want_log=True, want_state=True,
                              p_layout=L.P_DENSE if mode == L.MODE_DENSE else L.P_BLOCKS)
                # open-loop replay + single step
                f = torch.randn(B, T, m, dtype=torch.float64, device="cuda") + 100.0
                dq = 0.01 * torch.randn(B, T, 6, dtype=torch.float64, device="cuda")
                X = torch.zeros(B, N, dtype=torch.float64, device="cuda")
                E.replay(p, f, dq, f[:, 0].contiguous(), X, None, p_layout=L.P_BLOCKS if not kw else L.P_DENSE,
                         want_P=not kw)
        for method in (L.METHOD_ANALYTICAL, L.METHOD_KF, L.METHOD_MCKF, L.METHOD_GMCKF):
            p = E.filter_params(method, nf, mode=L.MODE_DENSE, kernel_bw=50.0, fpi_threshold=0.01, fpi_epoch_max=50)
            cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_WHITE, seed=2, std=1.0), T=T, dt=0.05, ibvs_gain=0.3,
                                    q_start=[0.0, -np.pi / 8, np.pi / 2 + np.pi / 8, 0.0, -np.pi / 2, 0.0],
                                    desired_f=DESIRED[nf], change_q_start=False)
            E.twist_loop(cfg, E.twist_cfg(x0_mode=L.X0_IMAGE_JACOBIAN, focal_x0=512.0, focal=512.0), 0, B, want_log=True,
                         want_state=True)
    # zero initial Jacobian: the pseudo-inverse meets a rank-deficient Gram matrix on the first frames, so the cooperative
    # one-sided Jacobi fallback (pinv_fallback_osj) runs in every family and for every M, with the per-frame trace on
    for nf in (4, 3, 1):
        m, N = 2 * nf, 12 * nf
        X0 = torch.zeros(B, N, dtype=torch.float64, device="cuda")
        for mode in (L.MODE_DENSE, L.MODE_BLOCKDIAG):
            p = E.filter_params(L.METHOD_KF, nf, mode=mode)
            cfg = E.closed_loop_cfg(p, noise, T=T, dt=0.05, ibvs_gain=0.2, q_start=Q_START, desired_f=DESIRED[nf])
            out = E.closed_loop(cfg, 0, B, X0_in=X0, want_trace=True,
                                dq_in=0.01 * torch.randn(B, T, 6, dtype=torch.float64, device="cuda"))
            assert int((out["flags"] & L.FLAG_RANK_DEFICIENT).ne(0).sum()) == B
        p = E.filter_params(L.METHOD_KF, nf, mode=L.MODE_DENSE)
        cfg = E.closed_loop_cfg(p, noise, T=T, dt=0.05, ibvs_gain=0.3, q_start=Q_START, desired_f=DESIRED[nf])
        out = E.twist_loop(cfg, E.twist_cfg(x0_mode=L.X0_ZEROS), 0, B)
        assert int((out["flags"] & L.FLAG_RANK_DEFICIENT).ne(0).sum()) == B
    q = torch.randn(B, 6, dtype=torch.float64, device="cuda")
    E.ur10_fk(q); E.project(4, q); E.initial_guess(4, q); E.noise_fill(noise, 0, B, T, 8); E.start_poses(Q_START, 5, True, 0, B)
    J = torch.randn(B, 8, 6, dtype=torch.float64, device="cuda")
    J[0] = 0.0
    E.ibvs(4, J, torch.randn(B, 8, dtype=torch.float64, device="cuda"), None, 0.2)
    st = torch.rand(B, 3, dtype=torch.float64, device="cuda")
    E.reduce_stats(st, torch.zeros(B, dtype=torch.int32, device="cuda"), 0, 4, 6)
    # log streaming: two streams, two device + two page-locked host buffers
    p = E.filter_params(L.METHOD_GMCKF, 4, mode=L.MODE_BLOCKDIAG, annealing=True)
    cfg = E.closed_loop_cfg(p, noise, T=T, dt=0.05, ibvs_gain=0.2, q_start=Q_START, desired_f=DESIRED[4])
    E.closed_loop_stream(cfg, 0, 50, chunk=16, sink=lambda *a: 0)
    torch.cuda.synchronize()
    print("sanitize.py: all launches done")


if __name__ == "__main__":
    main()
