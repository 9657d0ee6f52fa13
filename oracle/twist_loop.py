"""Oracle (test infrastructure): the camera-twist closed loop of the reference's tests/ scripts.

Restates the frame loop shared by /root/reference/tests/{ibvs,kalman,mckf}_{1,3,4}_feature(s).py
 on the kinematic plant of ``kinematics``:

    f_old = f ; f = detect(image) + noise                      kalman_4_features.py:113-128
    P = P + Q ; Z = f - f_old ; H = kron(I_m, dp^T)             :136-156   (H = 0 on the first frame)
    KF / MCKF update (simple-form or Joseph P)                  :159-161, mckf_4_features.py:156-195
    error = f - f* ; dp = -GAIN pinv(J) error [; dp[i] = 0]     :163-167, kalman_3_features.py:138-141
    dx = kron(I2, R06^T) dp ; dq = pinv(J_geom) dx              :168-172
    q <- q + dq TS                                              :173,192
    loop while norm(error) > ERROR_THRESHOLD                    :111

and the calibrated baseline (``ANALYTICAL``): J = the analytic interaction matrix of the noisy
features with the true camera-to-disc distance (ibvs_4_features.py:74-94) or the camera height
(ibvs_1_feature.py:55-70).  The regressor is the commanded camera twist dp, not dq; that is the
only structural difference from ``closed_loop.run_trial`` (experiment.py).

Pinned against the unmodified scripts executed with a stubbed simulator: oracle/make_golden_twist.py
-> tests/golden/ref_twist_*.npz, checked in tests/test_oracle_vs_golden.py.
"""
import numpy as np

from . import closed_loop, filters, kinematics, noise_model, stats as stats_mod

SUCCESS, FAIL = 0, 1
DEPTH_DISTANCE, DEPTH_HEIGHT = 0, 1          # computeZ (ur10_simulation.py:170-195) / getCameraHeight (:146-149)
X0_ZEROS, X0_IMAGE_JACOBIAN, X0_RANDOM = 0, 1, 2

DEFAULTS = dict(
    method=filters.KF, nf=4, dt=0.05, T=440, gain=0.5,
    q_start=[0.0, -np.pi / 8, np.pi / 2 + np.pi / 8, 0.0, -np.pi / 2, 0.0],        # kalman_4_features.py:31
    desired_f=[149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0],
    x0_mode=X0_IMAGE_JACOBIAN, focal_x0=1024.0, focal=kinematics.FOCAL, depth_mode=DEPTH_DISTANCE,
    err_threshold=-1.0, dp_zero_mask=0, square_inv=False,
    # True: R06 and J_geom of the START pose in every frame — kalman_4_features.py:165-169 and mckf_{1,3,4}_feature(s).py call
    # getCameraRotation() / jacobian() without recalculate_fkine, and UR10Simulation only refreshes its cached transforms
    # in fkine(recalculate=True) (ur10_simulation.py:97-110); the other five scripts refresh every frame
    stale_fk=False,
    kernel_bw=20.0, fpi_threshold=0.01, fpi_epoch_max=100, annealing=False, joseph=True, mckf_skip=False,
    noise_type=0, noise_params={}, hold_cnt=0,
    noise_seed=123456, pose_seed=12345, change_q_start=False, targets=None, P0=None, Q=None, R=None,
    v_max=0.0, a_max=0.0, servo_gain=0.0,
    sensor_offset=(0.0, 0.0, 0.0), disc_radius=0.0, servo_substeps=0, servo_kp=0.0, fov_check=False,      # scene plant
)


def make_cfg(**kw):
    cfg = dict(DEFAULTS)
    cfg.update(kw)
    return cfg


def image_jacobian(f, Zc, focal):
    """The analytic interaction matrix with RAW pixel coordinates (kalman_4_features.py:86-99,
    ibvs_4_features.py:82-94, experiment.py:147-159).  f [m], Zc [m/2] -> [m,6]."""
    m = len(f)
    J = np.zeros((m, 6))
    for i in range(m // 2):
        u, v = f[2 * i], f[2 * i + 1]
        J[2 * i, 0] = -focal / Zc[i]
        J[2 * i + 1, 1] = -focal / Zc[i]
        J[2 * i, 2] = u / Zc[i]
        J[2 * i + 1, 2] = v / Zc[i]
        J[2 * i, 3] = u * v / focal
        J[2 * i + 1, 3] = (focal ** 2 + v ** 2) / focal
        J[2 * i, 4] = -(focal ** 2 + u ** 2) / focal
        J[2 * i + 1, 4] = -u * v / focal
        J[2 * i, 5] = v
        J[2 * i + 1, 5] = -u
    return J


def depths(q, targets, depth_mode):
    if depth_mode == DEPTH_HEIGHT:
        return np.full(len(targets), kinematics.fkine(q)[2, 3])
    return kinematics.compute_Z(q, targets)


def twist_to_joint_rates(q, dp):
    """dx = kron(I2, R06^T) dp ; dq = pinv(J_geom) dx   (kalman_4_features.py:168-172); q = the pose the robot object's
    cached transforms were last refreshed at."""
    R = kinematics.fkine(q)[0:3, 0:3]
    dx = np.kron(np.eye(2), R.T) @ np.asarray(dp, dtype=float).reshape(6, 1)
    return (np.linalg.pinv(kinematics.jacobian(q)) @ dx).ravel()


def run_trial(cfg, trial=0, noise=None, q0=None, x0=None, record=False):
    nf = cfg["nf"]
    m, N = 2 * nf, 12 * nf
    dt, T = cfg["dt"], int(cfg["T"])
    targets = np.asarray(cfg["targets"]) if cfg["targets"] is not None else kinematics.TARGETS[nf]
    desired = np.asarray(cfg["desired_f"], dtype=float)[:m]
    method = cfg["method"]
    q = np.array(q0, dtype=float) if q0 is not None else \
        noise_model.start_pose(cfg["q_start"], cfg["pose_seed"], trial, cfg["change_q_start"])
    if noise is None:
        if cfg["noise_type"]:
            noise = noise_model.noise_stream(cfg["noise_type"], cfg["noise_params"], cfg["noise_seed"],
                                             trial, T, m, cfg["hold_cnt"])
        else:
            noise = np.zeros((T, m))
    P = np.eye(N) if cfg["P0"] is None else np.array(cfg["P0"], dtype=float)
    Q = np.eye(N) if cfg["Q"] is None else np.asarray(cfg["Q"], dtype=float)
    R = np.eye(m) if cfg["R"] is None else np.asarray(cfg["R"], dtype=float)

    # initial feature vector and state (kalman_4_features.py:40,74-101; kalman_1_feature.py:44,57; kalman_3:50)
    f = np.zeros(m)
    if cfg["x0_mode"] == X0_IMAGE_JACOBIAN:
        f = closed_loop.measure(cfg, q, targets)[0]
    if x0 is not None:
        X = np.array(x0, dtype=float).ravel()
    elif cfg["x0_mode"] == X0_IMAGE_JACOBIAN:
        X = image_jacobian(f, depths(q, targets, cfg["depth_mode"]), cfg["focal_x0"]).reshape(N)
    elif cfg["x0_mode"] == X0_RANDOM:
        X = noise_model.random_x0(cfg["noise_seed"], trial, N)
    else:
        X = np.zeros(N)

    q_log = np.zeros((T, 6)); f_log = np.zeros((T, m)); noise_log = np.zeros((T, m)); t_log = np.zeros(T)
    dp_log = np.zeros((T, 6)); dq_log = np.zeros((T, 6))
    X_log = np.zeros((T, N)) if record else None
    status = SUCCESS
    dp = np.zeros(6)
    v_prev = np.zeros(6)
    q_fk = q.copy()                                          # pose of the last fkine(recalculate=True): the start pose
    error = np.ones(m)                                       # kalman_4_features.py:70
    k = 0
    for k_it in range(T):
        if cfg["err_threshold"] >= 0.0 and not (np.linalg.norm(error) > cfg["err_threshold"]):
            break                                            # kalman_4_features.py:111
        f_old = f.copy()
        f_true, visible = closed_loop.measure(cfg, q, targets)
        if cfg.get("fov_check") and not np.all(visible):
            status = FAIL
            break
        f = f_true + noise[k_it]
        kappa = np.ones(m)
        if method == filters.ANALYTICAL:
            J = image_jacobian(f, depths(q, targets, cfg["depth_mode"]), cfg["focal"])
            X = J.reshape(N)
        else:
            Z = f - f_old
            h = np.zeros(6) if k_it == 0 else dp
            sigma = filters.annealed_bw(cfg["kernel_bw"], cfg["annealing"], k_it, T)
            X, P, K, kappa, info = filters.step_dense(
                method, X, P, h, Z, Q, R, sigma, cfg["fpi_threshold"], cfg["fpi_epoch_max"],
                cfg["joseph"], cfg["mckf_skip"], mckf_recompute_DW=True)
            J = X.reshape(m, 6)
        error = f - desired
        try:
            # kalman_3_features.py:138 / mckf_3_features.py:182 invert the square Jacobian with inv, not pinv
            Jinv = np.linalg.inv(J) if (cfg["square_inv"] and m == 6) else np.linalg.pinv(J)
            dp = (-cfg["gain"] * Jinv @ (kappa * error).reshape(m, 1)).ravel()
            for i in range(6):
                if cfg["dp_zero_mask"] >> i & 1:
                    dp[i] = 0.0                              # kalman_3_features.py:139-141
            dq = twist_to_joint_rates(q_fk if cfg.get("stale_fk") else q, dp)
        except np.linalg.LinAlgError:
            status = FAIL
            break
        if not (np.all(np.isfinite(dq)) and np.all(np.isfinite(X))):
            status = FAIL
            break
        q_log[k] = q; f_log[k] = f; noise_log[k] = noise[k_it]; t_log[k] = (k_it + 1) * dt
        dp_log[k] = dp; dq_log[k] = dq
        if record:
            X_log[k] = X
        k += 1
        if cfg.get("servo_substeps", 0) > 0:
            q, v_prev = closed_loop.servo_substepped(cfg, dq, q, v_prev)
        else:
            v_prev = closed_loop.servo_rate(cfg, dq, v_prev)
            q = q + v_prev * dt
    out = dict(status=status, frames=k, q_log=q_log[:k], f_log=f_log[:k], noise_log=noise_log[:k],
               t_log=t_log[:k], dp_log=dp_log[:k], dq_log=dq_log[:k], X=X, P=P, q_final=q)
    out["stats"] = stats_mod.trial_stats(out["t_log"], out["f_log"], desired)
    if record:
        out["X_log"] = X_log[:k]
    return out
