"""Oracle (test infrastructure): one closed-loop IBVS trial on the kinematic plant.

Follows /root/reference/experiment.py:48-359 (Experiment.run) with the
simulator replaced by ``kinematics`` (q <- q + dq*dt, pinhole projection of the
disc centres) and the time base fixed to T = int(t_max/dt) frames,
t_k = (k+1)*dt (SURVEY.md Appendix B 'Time base').  Same conventions as the
CUDA closed-loop kernel so trajectories are comparable frame by frame when
both consume the same noise stream.
"""
import numpy as np

from . import filters, ibvs, kinematics, noise_model, stats as stats_mod

SUCCESS, FAIL = 0, 1     # experiment.py:13-15

DEFAULTS = dict(
    method=filters.GMCKF, nf=4, dt=0.05, t_max=15.0, ibvs_gain=0.2,
    q_start=[0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0],
    desired_f=[149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0],
    initial_guess=True, kernel_bw=10.0, fpi_threshold=0.1, fpi_epoch_max=1000,
    annealing=False, joseph=True, mckf_skip=True,
    noise_type=noise_model.ALPHA_STABLE,
    noise_params=dict(alpha=1.5, beta=0.0, gamma=1.0, delta=0.0),
    hold_cnt=0, noise_seed=123456, pose_seed=12345, change_q_start=True,
    fov_check=False, P0=None, Q=None, R=None, targets=None, structured=False,
    v_max=0.0, a_max=0.0, servo_gain=0.0,        # joint servo model (SURVEY.md Appendix A 'Joints'); 0 = kinematic plant
    # scene plant (include/uvs_b200.h): sensor mount offset in DH-6, finite discs, sub-stepped P servo; all 0 = off
    sensor_offset=(0.0, 0.0, 0.0), disc_radius=0.0, servo_substeps=0, servo_kp=0.0,
)


def make_cfg(**kw):
    cfg = dict(DEFAULTS)
    cfg.update(kw)
    return cfg


def servo_rate(cfg, dq, v_prev):
    """Joint rates actually executed for the commanded dq under the scene's servo limits
    (include/uvs_b200.h: uvs_closed_loop_cfg.v_max / a_max): acceleration clamp, then velocity clamp."""
    v = np.array(dq, dtype=float)
    if cfg.get("servo_gain", 0.0) > 0.0:
        v = cfg["servo_gain"] * v
    if cfg.get("a_max", 0.0) > 0.0:
        dv = cfg["a_max"] * cfg["dt"]
        v = np.minimum(np.maximum(v, v_prev - dv), v_prev + dv)
    if cfg.get("v_max", 0.0) > 0.0:
        v = np.minimum(np.maximum(v, -cfg["v_max"]), cfg["v_max"])
    return v


def servo_substepped(cfg, dq, q, v_prev):
    """csrc/kernel_args.cuh servo_substepped: P controller towards q + dq dt at h = dt / substeps under a_max / v_max."""
    q = np.array(q, dtype=float)
    v_prev = np.array(v_prev, dtype=float)
    target = q + np.asarray(dq, dtype=float) * cfg["dt"]
    h = cfg["dt"] / cfg["servo_substeps"]
    dv = cfg.get("a_max", 0.0) * h
    for _ in range(cfg["servo_substeps"]):
        v = cfg["servo_kp"] * (target - q) / h
        if cfg.get("a_max", 0.0) > 0.0:
            v = np.minimum(np.maximum(v, v_prev - dv), v_prev + dv)
        if cfg.get("v_max", 0.0) > 0.0:
            v = np.minimum(np.maximum(v, -cfg["v_max"]), cfg["v_max"])
        q = q + v * h
        v_prev = v
    return q, v_prev


def scene_plant(cfg):
    return cfg.get("disc_radius", 0.0) > 0.0 or any(v != 0.0 for v in cfg.get("sensor_offset", (0, 0, 0))) or \
        cfg.get("servo_substeps", 0) > 0


def measure(cfg, q, targets):
    """Noise-free features and the 'visible' verdict under the configured plant."""
    if scene_plant(cfg):
        return kinematics.scene_features(q, targets, cfg.get("sensor_offset", (0, 0, 0)), cfg.get("disc_radius", 0.0),
                                         bool(cfg["fov_check"]))
    f = kinematics.project(q, targets)
    z = np.repeat(kinematics.depth(q, targets), 2)
    return f, (f >= 0.0) & (f < kinematics.RESOLUTION) & (z > 0.01) & np.isfinite(f)


def in_view(f, z):
    return bool(np.all(np.isfinite(f)) and np.all(f >= 0.0) and np.all(f < kinematics.RESOLUTION)
                and np.all(z > 0.01))


def run_trial(cfg, trial, noise=None, q0=None, x0=None, record=False, dq_in=None):
    """Returns dict(status, frames, stats[3], q_log, f_log, noise_log, t_log, ...).

    noise: optional [T,m] precomputed stream (else Philox from cfg);
    q0 / x0: optional overrides of the start pose / initial state;
    dq_in: optional [T,6] recorded command stream that replaces the computed command after every frame
    (uvs_closed_loop_io.dq_in: the replay gate of the closed-loop kernels); dq_log keeps the computed one.

    A frame on which the reference raises out of Experiment.run() — np.linalg.inv(Cx) with a zero weight
    (experiment.py:224), np.linalg.cholesky(P) (:203), a singular innovation covariance outside GMCKF's try
    block (:192, :238, :264) — ends the trial with FAIL (the CUDA kernels: csrc/common.cuh kAbortFlags).
    """
    nf = cfg["nf"]
    m, N = 2 * nf, 12 * nf
    dt = cfg["dt"]
    T = int(cfg["t_max"] / dt)
    targets = np.asarray(cfg["targets"]) if cfg["targets"] is not None else kinematics.TARGETS[nf]
    desired = np.asarray(cfg["desired_f"], dtype=float)[:m]
    method = cfg["method"]
    q = np.array(q0, dtype=float) if q0 is not None else \
        noise_model.start_pose(cfg["q_start"], cfg["pose_seed"], trial, cfg["change_q_start"])
    if noise is None:
        noise = noise_model.noise_stream(cfg["noise_type"], cfg["noise_params"], cfg["noise_seed"],
                                         trial, T, m, cfg["hold_cnt"])
    P = np.eye(N) if cfg["P0"] is None else np.array(cfg["P0"], dtype=float)
    Q = np.eye(N) if cfg["Q"] is None else np.asarray(cfg["Q"], dtype=float)
    R = np.eye(m) if cfg["R"] is None else np.asarray(cfg["R"], dtype=float)
    structured = cfg["structured"]
    if structured:
        Pb = filters.blocks_of(P, m)
        qd = filters.blocks_of(Q, m)
        rd = np.diag(R).copy()

    # experiment.py:86-117
    if x0 is not None:
        X = np.array(x0, dtype=float)
        f = measure(cfg, q, targets)[0] if cfg["initial_guess"] else np.zeros(m)
    elif cfg["initial_guess"]:
        X, f = kinematics.initial_guess(q, targets, measure(cfg, q, targets)[0] if scene_plant(cfg) else None)
    else:
        X = noise_model.random_x0(cfg["noise_seed"], trial, N)
        f = np.zeros(m)

    q_log = np.zeros((T, 6)); f_log = np.zeros((T, m)); noise_log = np.zeros((T, m))
    t_log = np.zeros(T); bw_log = np.zeros(T); dq_log = np.zeros((T, 6))
    X_log = np.zeros((T, N)) if record else None
    P_log = [] if record else None
    flags_log = np.zeros(T, dtype=np.int64); epoch_log = np.zeros(T, dtype=np.int64)
    status = SUCCESS
    abort_flags = 0
    dq = np.zeros(6)
    v_prev = np.zeros(6)
    k = 0
    for k_it in range(T):
        t = (k_it + 1) * dt
        f_old = f.copy()
        f_true, visible = measure(cfg, q, targets)
        if cfg["fov_check"] and not np.all(visible):
            status = FAIL          # NaN centroid -> pinv raises (utils.py:73-78, experiment.py:312-316)
            break
        f = f_true + noise[k_it]
        Z = f - f_old
        h = np.zeros(6) if k_it == 0 else dq         # experiment.py:183-188
        sigma = filters.annealed_bw(cfg["kernel_bw"], cfg["annealing"], k_it, T)
        try:
            if structured:
                X, Pb, K, kappa, info = filters.step_blockdiag(
                    method, X, Pb, h, Z, qd, rd, sigma, cfg["fpi_threshold"], cfg["fpi_epoch_max"],
                    cfg["joseph"], cfg["mckf_skip"])
            else:
                X, P, K, kappa, info = filters.step_dense(
                    method, X, P, h, Z, Q, R, sigma, cfg["fpi_threshold"], cfg["fpi_epoch_max"],
                    cfg["joseph"], cfg["mckf_skip"])
        except np.linalg.LinAlgError:            # cholesky(P) / inv(S): raised outside every try block of run()
            status = FAIL
            abort_flags = filters.FLAG_NOT_SPD
            break
        if info["flags"] & filters.FLAG_CX_SINGULAR:     # experiment.py:224 raises out of run()
            status = FAIL
            abort_flags = filters.FLAG_CX_SINGULAR
            break
        try:
            dq, _ = ibvs.ibvs_step(X, f, desired, kappa, cfg["ibvs_gain"])
        except np.linalg.LinAlgError:
            status = FAIL
            break
        if not (np.all(np.isfinite(dq)) and np.all(np.isfinite(X))):
            status = FAIL
            break
        q_log[k] = q; f_log[k] = f; noise_log[k] = noise[k_it]; t_log[k] = t
        bw_log[k] = sigma if method == filters.MCKF else -1      # experiment.py:330
        dq_log[k] = dq; flags_log[k] = info["flags"]; epoch_log[k] = info["epoch"]
        if record:
            X_log[k] = X
            P_log.append(Pb.copy() if structured else P.copy())
        k += 1
        if dq_in is not None:
            dq = np.asarray(dq_in[k_it], dtype=float)
        if cfg.get("servo_substeps", 0) > 0:
            q, v_prev = servo_substepped(cfg, dq, q, v_prev)
        else:
            v_prev = servo_rate(cfg, dq, v_prev)
            q = q + v_prev * dt                                   # experiment.py:320,335 (v = dq without limits)
    out = dict(status=status, frames=k, q_log=q_log[:k], f_log=f_log[:k], noise_log=noise_log[:k],
               t_log=t_log[:k], kernel_bw_log=bw_log[:k], dq_log=dq_log[:k],
               flags_log=flags_log[:k], epoch_log=epoch_log[:k], X=X, abort_flags=abort_flags,
               P=(Pb if structured else P))
    out["stats"] = stats_mod.trial_stats(out["t_log"], out["f_log"], desired)
    if record:
        out["X_log"] = X_log[:k]
        out["P_log"] = np.array(P_log)
    return out


def replay(method, nf, f_stream, dq_stream, f_init, X0, P0=None, Q=None, R=None, kernel_bw=10.0,
           annealing=False, k_max=None, fpi_threshold=0.1, fpi_epoch_max=1000, joseph=True,
           mckf_skip=True, structured=False):
    """Open-loop replay: host-supplied noisy features f[T,m] and the recorded
    regressor stream dq[T,6] (dq[k] was produced at frame k and is frame k+1's
    regressor).  Returns X[T,N], P[T,...], kappa[T,m], flags[T], epochs[T]."""
    m, N = 2 * nf, 12 * nf
    T = len(f_stream)
    k_max = T if k_max is None else k_max
    P = np.eye(N) if P0 is None else np.array(P0, dtype=float)
    Q = np.eye(N) if Q is None else np.asarray(Q, dtype=float)
    R = np.eye(m) if R is None else np.asarray(R, dtype=float)
    if structured:
        Pb = filters.blocks_of(P, m); qd = filters.blocks_of(Q, m); rd = np.diag(R).copy()
    X = np.array(X0, dtype=float)
    f = np.array(f_init, dtype=float)
    Xs = np.zeros((T, N)); Ps = []; kap = np.zeros((T, m))
    flags = np.zeros(T, dtype=np.int64); epochs = np.zeros(T, dtype=np.int64)
    for k in range(T):
        f_old, f = f, np.asarray(f_stream[k], dtype=float)
        Z = f - f_old
        h = np.zeros(6) if k == 0 else np.asarray(dq_stream[k - 1], dtype=float)
        sigma = filters.annealed_bw(kernel_bw, annealing, k, k_max)
        if structured:
            X, Pb, K, kappa, info = filters.step_blockdiag(method, X, Pb, h, Z, qd, rd, sigma,
                                                           fpi_threshold, fpi_epoch_max, joseph, mckf_skip)
            Ps.append(Pb.copy())
        else:
            X, P, K, kappa, info = filters.step_dense(method, X, P, h, Z, Q, R, sigma,
                                                      fpi_threshold, fpi_epoch_max, joseph, mckf_skip)
            Ps.append(P.copy())
        Xs[k] = X; kap[k] = kappa; flags[k] = info["flags"]; epochs[k] = info["epoch"]
    return Xs, np.array(Ps), kap, flags, epochs
