"""GPU tests of the batch driver (main.py equivalent) and of the BASELINE configurations that are not the
bench line: configs[1] (MCKF, 3 features, 10k trials x 500 frames, open-loop replay) and configs[4] (noise
x estimator x feature-count sweep).  Everything goes through the C ABI."""
import json

import numpy as np
import pytest

from conftest import rel_err
from test_sweep_host import BASE_CONFIG, REFERENCE_COLUMNS

pytestmark = pytest.mark.gpu

TOL_STEP = 1e-9


@pytest.fixture(scope="module")
def uvs():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import uvs_b200
    uvs_b200._lib.load()
    return uvs_b200


def _cfg(method="GMCKF", annealing=True, **noise):
    cfg = json.loads(json.dumps(BASE_CONFIG))
    cfg["estimator"]["method"] = method
    cfg["estimator"]["estimator_params"]["annealing"] = annealing
    cfg["noise"].update(noise)
    return cfg


def test_run_sweep_is_the_main_py_queue(uvs):
    """12 levels x epoch experiments in one launch: trial k has level k // epoch, the start pose / noise streams
    of global trial k, and the per-level summary equals the oracle statistics of the per-trial table."""
    from oracle import closed_loop as CL, stats as ostats
    S = uvs.sweep
    epoch = 6
    res = S.run_sweep(_cfg(), epoch=epoch, want_trials=True)
    assert res["n_trials"] == 12 * epoch and np.allclose(res["levels"], np.linspace(1, 2, 12))
    st, status = res["stats"], res["status"]
    for lv in range(12):
        sl = slice(lv * epoch, (lv + 1) * epoch)
        ref = ostats.config_stats(st[sl], status[sl])
        row = res["summary"][lv]
        assert row["n_ok"] == ref["n_ok"] and row["n_fail"] == ref["n_fail"]
        if ref["n_ok"] > 1:
            assert np.isclose(row["itae_mean"], ref["itae_mean"], rtol=1e-12)
            assert np.isclose(row["itae_std"], ref["itae_std"], rtol=1e-8)
            assert np.isclose(res["quantiles"][lv, 2, 1], ref["itae_median"], rtol=1e-12)
    # three experiments of different levels against the oracle's closed loop on the same Philox streams
    for k in (1, 3 * epoch + 2, 11 * epoch + 5):
        ocfg = CL.make_cfg(annealing=True, noise_params=dict(alpha=float(res["levels"][k // epoch]), beta=0.0,
                                                             gamma=1.0, delta=0.0))
        ref = CL.run_trial(ocfg, k)
        assert ref["status"] == status[k]
        assert np.max(np.abs(st[k] - ref["stats"]) / np.abs(ref["stats"])) < 1e-6, k
    # the dense family gives the same table (closed loop: chaotic beyond ~1e-6 on the odd trial)
    res_d = S.run_sweep(_cfg(), epoch=epoch, want_trials=True, mode=uvs._lib.MODE_DENSE)
    ok = (status == 0) & (res_d["status"] == 0)
    assert np.median(np.abs(res_d["stats"][ok] - st[ok]) / st[ok]) < 1e-9


def test_results_csv_has_the_reference_schema(uvs, tmp_path):
    """results.csv: main.py:152-196 column for column; one row per frame; experiment constants repeated."""
    import pandas as pd
    S = uvs.sweep
    path = tmp_path / "results.csv"
    res = S.run_sweep(_cfg("MCKF", annealing=False), epoch=2, csv_path=str(path), csv_trials=1, want_trials=True)
    df = pd.read_csv(path)
    assert list(df.columns) == REFERENCE_COLUMNS
    ids = sorted(df["experiment_id"].unique())
    assert ids == [2 * lv for lv in range(12)]                       # first experiment of every level
    g = df[df["experiment_id"] == 4]
    assert len(g) == 300 and np.allclose(g["t"].values, (np.arange(300) + 1) * 0.05)
    assert (g["status"] == "ExperimentStatus.SUCCESS").all() or (g["status"] == "ExperimentStatus.FAIL").all()
    assert np.allclose(g["rho"].values, res["levels"][2])           # the column called rho holds alpha (main.py:155)
    assert (g["kernel_bw"] == 10.0).all()                            # MCKF logs its bandwidth, experiment.py:330
    assert (g["desired_f_3"] == 125.0).all()
    # the logged rows are the trajectory the statistics were computed from
    f = g[[f"f_{i + 1}" for i in range(8)]].values
    des = np.array(BASE_CONFIG["experiments"]["desired_f"])
    itae = np.linalg.norm(g["t"].values @ np.abs(des[None] - f))
    assert np.isclose(itae, res["stats"][4, 2], rtol=1e-9)
    # f = projection(q) + noise, row by row
    q = g[[f"q_{i + 1}" for i in range(6)]].values
    proj, _ = uvs.engine.project(4, uvs.engine.dev(q))
    nz = g[[f"noise_{i + 1}" for i in range(8)]].values
    assert np.allclose(proj.cpu().numpy() + nz, f, rtol=0, atol=1e-9)
    # GMCKF logs -1 (experiment.py:330)
    path2 = tmp_path / "r2.csv"
    S.run_sweep(_cfg("GMCKF"), epoch=1, csv_path=str(path2), csv_trials=None)
    df2 = pd.read_csv(path2)
    assert (df2["kernel_bw"] == -1.0).all() and sorted(df2["experiment_id"].unique()) == list(range(12))


def test_device_and_host_level_reductions_agree(uvs):
    import torch
    S, E, L = uvs.sweep, uvs.engine, uvs._lib
    rng = np.random.default_rng(5)
    B, epoch, trial0 = 5000, 700, 350
    stats = rng.lognormal(9, 1, size=(B, 3))
    status = (rng.uniform(size=B) < 0.2).astype(np.int32)
    sums, n_fail = E.reduce_stats(E.dev(stats), E.dev(status, dtype=torch.int32), trial0, epoch, 12)
    hs, hf = S.level_sums_host(stats, status, trial0, epoch, 12)
    assert torch.allclose(sums.cpu(), hs, rtol=1e-12) and torch.equal(n_fail.cpu(), hf)


# ------------------------------------------------------------------ BASELINE configs[1]
def _smooth_streams(uvs, B, T, nf, seed, noise_std=1.0):
    """per-trial smooth joint trajectories q_k = q0 + 0.1 sin(2 pi k / T + phi_j), noisy
    projected features f_k and the regressor stream dq_k = (q_{k+1} - q_k) / dt (frame k+1 regresses on dq_k)."""
    rng = np.random.default_rng(seed)
    q0 = np.array([0.0, -np.pi / 8, 5 * np.pi / 8, 0.0, -np.pi / 2, 0.0])
    phi = rng.uniform(0, 2 * np.pi, size=(B, 1, 6))
    k = np.arange(T + 2)[None, :, None]
    q = q0[None, None, :] + 0.1 * np.sin(2 * np.pi * k / T + phi)          # q[0] initial pose, q[k+1] pose at frame k
    f_true, _ = uvs.engine.project(nf, uvs.engine.dev(q[:, :T + 1].reshape(-1, 6)))
    f_true = f_true.cpu().numpy().reshape(B, T + 1, 2 * nf)
    f = f_true + noise_std * rng.standard_normal(f_true.shape)
    dq = (q[:, 2:] - q[:, 1:-1]) / 0.05                                     # dq[k]: the command issued at frame k
    assert dq.shape == (B, T, 6)
    return np.ascontiguousarray(f[:, 1:]), np.ascontiguousarray(dq), np.ascontiguousarray(f[:, 0])


@pytest.mark.parametrize("kernel_bw", [10.0, 200.0])
def test_config1_mckf_3_features_replay(uvs, kernel_bw):
    """MCKF, 3 features (36 states), WHITE noise std 1, 10 000 trials x 500 frames, open-loop replay.
    A 24-trial subset is checked against the oracle every frame (X) and on the last frame (P) to 1e-9; the full
    batch is checked through split invariance, family agreement and determinism."""
    import torch
    from oracle import closed_loop as CL, filters as F
    L, E = uvs._lib, uvs.engine
    B, T, nf, sub = 10_000, 500, 3, 24
    N = 12 * nf
    f, dq, f0 = _smooth_streams(uvs, B, T, nf, seed=12345)
    rng = np.random.default_rng(7)
    X0 = rng.uniform(0, 1, size=(B, N))                    # tests/kalman_3_features.py: X0 = rand
    fd, dqd, f0d = E.dev(f), E.dev(dq), E.dev(f0)
    outs = {}
    for mode, layout in ((L.MODE_BLOCKDIAG, L.P_BLOCKS), (L.MODE_DENSE, L.P_BLOCKS)):
        p = E.filter_params(3, nf, mode=mode, kernel_bw=kernel_bw, fpi_threshold=0.1, fpi_epoch_max=1000)
        X = E.dev(X0)
        P = torch.eye(6, dtype=torch.float64, device="cuda").repeat(B, 2 * nf, 1, 1).contiguous()
        want_X = mode == L.MODE_BLOCKDIAG
        if mode == L.MODE_DENSE:                           # dense family on the first 2048 trials
            X, P = X[:2048].contiguous(), P[:2048].contiguous()
            X_out, _, _, flags = E.replay(p, fd[:2048], dqd[:2048], f0d[:2048], X, P, p_layout=layout, want_X=False)
        else:
            X_out, _, _, flags = E.replay(p, fd, dqd, f0d, X, P, p_layout=layout, want_X=False)
        torch.cuda.synchronize()
        outs[mode] = (X.cpu().numpy(), P.cpu().numpy(), flags.cpu().numpy())
    Xb, Pb, fb = outs[L.MODE_BLOCKDIAG]
    Xd, Pd, fdn = outs[L.MODE_DENSE]
    assert np.isfinite(Xb).all() and np.isfinite(Pb).all()
    # per-frame parity on the subset
    p = E.filter_params(3, nf, mode=L.MODE_BLOCKDIAG, kernel_bw=kernel_bw)
    Xs = E.dev(X0[:sub])
    X_out, P_out, _, flags = E.replay(p, fd[:sub], dqd[:sub], f0d[:sub], Xs, None, p_layout=L.P_BLOCKS, want_P=True)
    X_out, P_out, flags = X_out.cpu().numpy(), P_out.cpu().numpy(), flags.cpu().numpy()
    for b in range(sub):
        Xr, Pr, _, fl, ep = CL.replay(F.MCKF, nf, f[b], dq[b], f0[b], X0[b], kernel_bw=kernel_bw, structured=True)
        err = np.linalg.norm(X_out[b] - Xr, axis=1) / np.linalg.norm(Xr, axis=1)
        assert err.max() <= TOL_STEP, (b, err.max())
        assert rel_err(P_out[b, -1], Pr[-1]) <= TOL_STEP
        assert np.array_equal(flags[b] >> L.FLAG_EPOCH_SHIFT, ep)
        # the subset run and the 10k run are the same trials
        assert np.array_equal(X_out[b, -1], Xb[b]) and np.array_equal(P_out[b, -1], Pb[b])
    # families agree on the final state of 2048 trials (fixed-point counts are identical, so no chaos here)
    assert np.array_equal(fdn >> L.FLAG_EPOCH_SHIFT, fb[:2048] >> L.FLAG_EPOCH_SHIFT)
    e = np.linalg.norm(Xd - Xb[:2048], axis=1) / np.linalg.norm(Xb[:2048], axis=1)
    assert e.max() <= 1e-9
    assert rel_err(Pd, Pb[:2048]) <= 1e-9


# ------------------------------------------------------------------ BASELINE configs[4]
NOISES = {
    "WHITE_NOISE": dict(std=1.0),
    "GAUSSIAN_MIXTURE": dict(std=1.0, mean=50.0, rho=0.1),
    "GAUSSIAN_BIMODAL": dict(std=1.0, mean=50.0, rho=0.1),
    "ALPHA_STABLE": dict(alpha=1.3, beta=0.0, gamma=1.0, delta=0.0),
}
DESIRED = {1: [128.0, 128.0], 3: [153.5, 144.5, 125.5, 116.5, 99.5, 144.5],
           4: [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0]}


@pytest.mark.parametrize("nf", [1, 3, 4])
def test_config4_noise_estimator_sweep(uvs, nf):
    """{WHITE, MIXTURE, BIMODAL, ALPHA_STABLE} x kernel_bw {1, 10, 20, 200} + annealing x random starts x
    {KF, MCKF, RMCKF} at 1/3/4 features: every cell runs, both families agree, and under impulsive noise the
    robust estimator with annealing does not do worse than KF (the paper's claim, results1.fig)."""
    L, E = uvs._lib, uvs.engine
    B = 256
    q_start = [0.0, -np.pi / 8, 5 * np.pi / 8, 0.0, -np.pi / 2, 0.0]
    med = {}
    for nname, npar in NOISES.items():
        prof = uvs.NoiseProfiler(2 * nf, uvs.NoiseType[nname], seed=99, noise_params=npar)
        for method in ("KF", "MCKF", "RMCKF"):
            for bw, ann in ((1.0, True), (10.0, True), (20.0, False), (200.0, False)):
                if method == "KF" and bw != 10.0:
                    continue
                runs = {}
                nd = 32 if nf != 3 else B
                for mode in (L.MODE_BLOCKDIAG, L.MODE_DENSE):
                    exp = uvs.BatchExperiment(q_start, DESIRED[nf], prof, 0.05, 15, 0.2, uvs.Method[method],
                                              n_trials=B if mode == L.MODE_BLOCKDIAG else nd, trial0=0,
                                              method_params=dict(initial_guess=True, kernel_bw=bw, fpi_threshold=0.1,
                                                                 fpi_epoch_max=1000, annealing=ann, mode=mode))
                    runs[mode] = exp.run()
                a, d = runs[L.MODE_BLOCKDIAG], runs[L.MODE_DENSE]
                ok = a["status"] == 0
                assert np.isfinite(a["stats"][ok]).all()
                both = (a["status"][:nd] == 0) & (d["status"] == 0)
                assert (a["status"][:nd] == d["status"]).mean() >= 0.9
                if nf != 3:
                    if both.any():
                        rel = np.abs(a["stats"][:nd][both] - d["stats"][both]) / np.abs(d["stats"][both])
                        assert np.median(rel) < 1e-8, (nname, method, bw, np.median(rel))
                else:
                    # 3 points give a square 6x6 image Jacobian that passes close to its singular configurations:
                    # the closed loop amplifies rounding differences within ~10 frames (the oracle diverges from
                    # either family in the same way), so the two families are compared as distributions
                    ma, md = np.median(a["stats"][ok, 2]), np.median(d["stats"][d["status"] == 0, 2])
                    assert abs(np.log(ma / md)) < 0.7, (nname, method, bw, ma, md)
                med[(nname, method, bw)] = (np.median(a["stats"][ok, 2]) if ok.any() else np.inf, ok.mean())
    if nf == 4:
        for nname in ("GAUSSIAN_MIXTURE", "ALPHA_STABLE"):
            assert med[(nname, "RMCKF", 10.0)][0] <= 1.05 * med[(nname, "KF", 10.0)][0], (nname, med)
    assert all(v[1] > 0.5 for v in med.values()), med


# ------------------------------------------------------------------ empty / ragged / degenerate inputs
@pytest.mark.parametrize("mode", ["dense", "blockdiag"])
def test_empty_ragged_and_degenerate_batches(uvs, mode):
    """B = 0 is a no-op, ragged batch sizes (not a multiple of the warps per CTA / trials per warp) return exactly
    the rows of a larger launch, T = 1 runs, T = 0 and unknown sizes are argument errors, a NaN start state is a
    FAIL (experiment.py:313-316) that does not disturb its neighbours."""
    import torch
    L, E = uvs._lib, uvs.engine
    md = L.MODE_DENSE if mode == "dense" else L.MODE_BLOCKDIAG
    q0 = [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0]
    des = [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0]

    def cfg(T=300, nf=4, method=5):
        p = E.filter_params(method, nf, mode=md, annealing=True)
        return E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=7, alpha=1.4), T=T, dt=0.05, ibvs_gain=0.2,
                                 q_start=q0, desired_f=des[:2 * nf] if nf == 4 else DESIRED[nf])

    out0 = E.closed_loop(cfg(), 0, 0)
    assert out0["stats"].shape == (0, 3)
    big = E.closed_loop(cfg(), 100, 23)
    torch.cuda.synchronize()
    bs, bst = big["stats"].cpu().numpy(), big["status"].cpu().numpy()
    for off, n in ((0, 1), (3, 5), (16, 7), (0, 23)):
        part = E.closed_loop(cfg(), 100 + off, n)
        assert np.array_equal(part["stats"].cpu().numpy(), bs[off:off + n])
        assert np.array_equal(part["status"].cpu().numpy(), bst[off:off + n])
    one = E.closed_loop(cfg(T=1), 0, 3)
    assert (one["frames"].cpu().numpy() == 1).all() and np.isfinite(one["stats"].cpu().numpy()).all()
    with pytest.raises(L.UvsError):
        E.closed_loop(cfg(T=0), 0, 3)
    # NaN in one trial's start state: that trial fails, the others are bit-identical to a clean run
    X0, _ = E.initial_guess(4, E.start_poses(q0, 12345, True, 100, 23))
    clean = E.closed_loop(cfg(), 100, 23, X0_in=X0)
    X0n = X0.clone()
    X0n[4, 7] = float("nan")
    dirty = E.closed_loop(cfg(), 100, 23, X0_in=X0n)
    ds, dst = dirty["stats"].cpu().numpy(), dirty["status"].cpu().numpy()
    assert dst[4] == 1 and (dirty["flags"].cpu().numpy()[4] & L.FLAG_NONFINITE)
    keep = np.arange(23) != 4
    assert np.array_equal(ds[keep], clean["stats"].cpu().numpy()[keep])
    assert np.array_equal(dst[keep], clean["status"].cpu().numpy()[keep])
    # replay: ragged batch, T = 1; zero-length batch
    rng = np.random.default_rng(0)
    for B in (1, 5, 9):
        f = E.dev(rng.standard_normal((B, 1, 8)) + 128)
        dq = E.dev(rng.standard_normal((B, 1, 6)) * 0.01)
        X = E.dev(rng.uniform(size=(B, 48)))
        p = E.filter_params(2, 4, mode=md)
        Xo, _, _, fl = E.replay(p, f, dq, E.dev(f[:, 0] * 0 + 128), X, None, p_layout=L.P_BLOCKS)
        assert Xo.shape == (B, 1, 48) and torch.isfinite(Xo).all()
        assert torch.equal(Xo[:, 0], X)                          # first frame: H = 0, the state does not move
    with pytest.raises(L.UvsError):                              # nf must be 1, 3 or 4
        E.replay(E.filter_params(2, 2, mode=md), E.dev(np.zeros((1, 1, 4))), E.dev(np.zeros((1, 1, 6))),
                 E.dev(np.zeros((1, 4))), E.dev(np.zeros((1, 24))), None, p_layout=L.P_BLOCKS)
    # empty batches of the other entry points are no-ops too
    z = torch.empty((0, 6), dtype=torch.float64, device="cuda")
    assert E.ur10_fk(z)[0].shape == (0, 4, 4) and E.project(4, z)[0].shape == (0, 8)
    assert E.noise_fill(E.noise_params(L.NOISE_WHITE, seed=1, std=1.0), 0, 0, 10, 8).shape == (0, 10, 8)


# ------------------------------------------------------------------ trajectory-log streaming
def test_stream_matches_the_one_shot_log(uvs):
    """uvs_closed_loop_stream: chunked launches, two device + two page-locked host buffers, the copy of a chunk overlapping
    the next launch.  Every chunk handed to the sink equals the log / statistics of one big logged launch, ragged last
    chunk included; a sink that asks to stop ends the stream with an error."""
    L, E = uvs._lib, uvs.engine
    p = E.filter_params(5, 4, mode=L.MODE_BLOCKDIAG, annealing=True)
    cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=11, alpha=1.4), T=300, dt=0.05, ibvs_gain=0.2,
                            q_start=[0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0], desired_f=DESIRED[4])
    B, trial0 = 1000, 77
    ref = E.closed_loop(cfg, trial0, B, want_log=True)
    rlog, rst, rss, rfr = (ref[k].cpu().numpy() for k in ("log", "stats", "status", "frames"))
    got = {"log": np.zeros_like(rlog), "stats": np.zeros_like(rst), "chunks": []}

    def sink(first, log, st, ss, fr):
        a = first - trial0
        got["log"][a:a + len(st)] = log
        got["stats"][a:a + len(st)] = st
        got["chunks"].append((first, len(st)))
        assert np.array_equal(ss, rss[a:a + len(st)]) and np.array_equal(fr, rfr[a:a + len(st)])

    info = E.closed_loop_stream(cfg, trial0, B, chunk=192, sink=sink)
    assert got["chunks"] == [(trial0 + 192 * i, min(192, B - 192 * i)) for i in range(6)]      # in order, ragged tail
    assert np.array_equal(got["log"], rlog) and np.array_equal(got["stats"], rst)
    assert info["d2h_gbs"] > 0.5 and info["wall_s"] > 0
    with pytest.raises(L.UvsError):
        E.closed_loop_stream(cfg, trial0, B, chunk=192, sink=lambda *a: True)
    with pytest.raises(ZeroDivisionError):                                      # a sink's exception resurfaces in Python
        E.closed_loop_stream(cfg, trial0, B, chunk=192, sink=lambda *a: 1 / 0)
    assert E.closed_loop_stream(cfg, 0, 0)["wall_s"] == 0.0                     # empty batch


def test_run_sweep_streams_every_experiment_to_csv(uvs, tmp_path):
    """main.py:152-196 logs every experiment: csv_trials=None streams the whole sweep once (statistics and rows from the
    same launches) and writes the rows with the vectorised writer; same table as the per-level logged launches."""
    import pandas as pd
    S = uvs.sweep
    a, b = tmp_path / "stream.csv", tmp_path / "levels.csv"
    res = S.run_sweep(_cfg("GMCKF"), epoch=3, csv_path=str(a), csv_trials=None, want_trials=True, stream_chunk=7)
    res2 = S.run_sweep(_cfg("GMCKF"), epoch=3, csv_path=str(b), csv_trials=3, want_trials=True)
    assert np.array_equal(res["stats"], res2["stats"]) and res["stream"]["rows"] == int(36 * 300)
    da, db = pd.read_csv(a), pd.read_csv(b)
    assert list(da.columns) == REFERENCE_COLUMNS and len(da) == len(db) == 36 * 300
    assert da.equals(db)
    assert sorted(da["experiment_id"].unique()) == list(range(36))
