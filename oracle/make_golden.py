"""Oracle (test infrastructure): mint golden vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden            # writes tests/golden/*.npz

It imports /root/reference/experiment.py and noise.py as they are, drives
``Experiment.run()`` with ``fake_robot.FakeRobot`` (kinematic plant, features
passed through a patched ``detect4Circles``), and snapshots the reference's own
frame locals (X, P, K, dq, ...) once per frame.  Nothing from the reference is
copied into the repo; only its numerical outputs are stored.

Files written (all float64 unless noted):
  ref_run_<name>.npz     closed-loop run of one estimator: config, q_start, f_init, X0, the
                         noise / f / q / dq logs, X every frame, the diagonal 6x6 blocks of P on
                         ``p_frames``, max |off-block P|, kappa (GMCKF), kernel_bw, epoch, skip.
  ref_dense_<name>.npz   same with general SPD P0/Q/R injected through a numpy proxy that
                         answers the reference's first three ``np.eye`` calls
                         (experiment.py:73,75,76); full P stored on ``p_frames``.
  ref_stats_<name>.npz   ISE/IAE/ITAE of 256 reference closed-loop trials (random start poses).
  ref_noise.npz          10 000 reference NoiseProfiler draws per distribution (for KS tests) and
                         a hold-logic trace produced with injected generators.
"""
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from . import kinematics, noise_model, stats as stats_mod       # noqa: E402
from .fake_robot import FakeRobot                                # noqa: E402


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; goldens can only be minted in the build container")
    sys.path.insert(0, REF)
    import experiment as ref_experiment      # noqa
    import noise as ref_noise                # noqa
    ref_experiment.detect4Circles = lambda image: np.array(image, dtype=float).copy()
    return ref_experiment, ref_noise


class _NumpyProxy:
    """Answers the first len(queue) ``eye`` calls from ``queue``; everything
    else is numpy.  Lets the unmodified run() start from general P0, Q, R."""

    def __init__(self, queue):
        self._queue = list(queue)

    def eye(self, *a, **kw):
        if self._queue:
            return np.array(self._queue.pop(0), dtype=float)
        return np.eye(*a, **kw)

    def __getattr__(self, name):
        return getattr(np, name)


def spd(n, seed, scale=1.0, coupling=0.3):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, n))
    M = np.eye(n) + coupling * (A @ A.T) / n
    return scale * (M + M.T) / 2


P_FRAMES_300 = sorted(set(list(range(20)) + list(range(20, 300, 10)) + [299]))


def run_reference(ref_experiment, ref_noise, method_name, noise_type_name, noise_params, q_start,
                  method_params, hold=False, hold_cnt=0, noise_seed=123456, dt=0.05, t_max=15.0,
                  ibvs_gain=0.2, fov_check=False, eye_queue=None,
                  desired_f=(149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0)):
    Method = ref_experiment.Method
    NoiseType = ref_noise.NoiseType
    prof = ref_noise.NoiseProfiler(8, NoiseType[noise_type_name], seed=noise_seed, noise_hold=hold,
                                   noise_hold_cnt=hold_cnt, noise_params=dict(noise_params))
    robot = FakeRobot(dt=dt, t_max=t_max, nf=4, fov_check=fov_check)
    exp = ref_experiment.Experiment(q_start=np.array(q_start, dtype=float), desired_f=list(desired_f),
                                    noise_prof=prof, t_s=dt, t_max=t_max, ibvs_gain=ibvs_gain, robot=robot,
                                    method=Method[method_name], method_params=dict(method_params))
    saved_np = ref_experiment.np
    if eye_queue is not None:
        ref_experiment.np = _NumpyProxy(eye_queue)
    try:
        status, t_log, error_log, q_log, f_log, desired_log, camera_log, noise_log, bw_log = exp.run()
    finally:
        ref_experiment.np = saved_np
    return dict(status=status.value, t_log=t_log, error_log=error_log, q_log=q_log, f_log=f_log,
                camera_log=camera_log, noise_log=noise_log, kernel_bw_log=bw_log, snaps=robot.snapshots)


def pack_run(res, q_start, dense, p_frames, extra):
    snaps = res["snaps"]
    T = len(snaps)
    X = np.stack([s["X"].ravel() for s in snaps])
    dq = np.stack([s["dq"].ravel() for s in snaps])
    p_frames = [k for k in p_frames if k < T]
    off = max(_off_block(s["P"]) for s in snaps)
    if dense:
        P = np.stack([snaps[k]["P"] for k in p_frames])
    else:
        P = np.stack([np.stack([snaps[k]["P"][6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(8)])
                      for k in p_frames])
    X0, f_init = kinematics.initial_guess(np.asarray(q_start, dtype=float), kinematics.TARGETS[4])
    out = dict(status=res["status"], frames=T, q_start=np.asarray(q_start, dtype=float), f_init=f_init,
               X0=X0, t_log=res["t_log"], q_log=res["q_log"], f_log=res["f_log"], noise_log=res["noise_log"],
               camera_log=res["camera_log"], kernel_bw_log=res["kernel_bw_log"], dq=dq, X=X, P=P,
               p_frames=np.array(p_frames), off_block_max=off,
               epoch=np.array([s.get("epoch", 0) for s in snaps]),
               skip=np.array([int(bool(s.get("skip_correction", False))) for s in snaps]),
               kernel_bw=np.array([float(s.get("kernel_bw", -1.0)) for s in snaps]),
               kappa=np.stack([np.asarray(s["kappa"], dtype=float).ravel() for s in snaps]))
    out.update(extra)
    return out


def _off_block(P):
    Pc = P.copy()
    for i in range(8):
        Pc[6 * i:6 * i + 6, 6 * i:6 * i + 6] = 0
    return float(np.abs(Pc).max())


class _InjectedGen:
    """Replaces a numpy Generator inside the reference NoiseProfiler so that its
    draws come from a known list (used to pin the hold logic exactly)."""

    def __init__(self, values, clock):
        self.values = values
        self.clock = clock

    def normal(self, loc=0.0, scale=1.0):
        return self.values[self.clock["k"]]

    def uniform(self, low=0.0, high=1.0):
        return self.values[self.clock["k"]]


def main():
    ref_experiment, ref_noise = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    q_nom = [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0]
    q_a = [-0.12, -0.35, 1.96349541, 0.0, -1.57079633, 0.0]
    q_b = [-0.25, -0.55, 1.96349541, 0.0, -1.57079633, 0.0]
    mp = dict(initial_guess=True, kernel_bw=10, fpi_threshold=0.1, fpi_epoch_max=1000, annealing=False)

    runs = {
        "kf_white": ("KF", "WHITE_NOISE", dict(std=1.0), q_a, dict(mp), False, 0),
        "gmckf_alpha15_anneal": ("GMCKF", "ALPHA_STABLE", dict(alpha=1.5, beta=0, gamma=1, delta=0), q_a,
                                 dict(mp, annealing=True), False, 0),
        "gmckf_mixture_hold": ("GMCKF", "GAUSSIAN_MIXTURE", dict(std=1.0, mean=50.0, rho=0.1), q_b,
                               dict(mp), True, 10),
        "mckf_alpha12": ("MCKF", "ALPHA_STABLE", dict(alpha=1.2, beta=0, gamma=1, delta=0), q_a,
                         dict(mp), False, 0),
        "mckf_white_anneal": ("MCKF", "WHITE_NOISE", dict(std=1.0), q_b, dict(mp, annealing=True), False, 0),
        "imcckf_bimodal_anneal": ("IMCCKF", "GAUSSIAN_BIMODAL", dict(std=1.0, mean=50.0, rho=0.15), q_a,
                                  dict(mp, annealing=True), False, 0),
        "kf_nominal_fov": ("KF", "WHITE_NOISE", dict(std=1.0), q_nom, dict(mp), False, 0),
    }
    for name, (method, ntype, nparams, q0, mparams, hold, hold_cnt) in runs.items():
        fov = name.endswith("_fov")
        res = run_reference(ref_experiment, ref_noise, method, ntype, nparams, q0, mparams, hold, hold_cnt,
                            fov_check=fov)
        extra = dict(method=method, noise_type=ntype, annealing=int(mparams["annealing"]),
                     kernel_bw0=float(mparams["kernel_bw"]), fpi_threshold=mparams["fpi_threshold"],
                     fpi_epoch_max=mparams["fpi_epoch_max"], hold_cnt=hold_cnt if hold else 0,
                     fov_check=int(fov), **{"np_" + k: float(v) for k, v in nparams.items()})
        if res["snaps"]:
            np.savez_compressed(os.path.join(OUT, f"ref_run_{name}.npz"),
                                **pack_run(res, q0, False, P_FRAMES_300, extra))
        else:
            np.savez_compressed(os.path.join(OUT, f"ref_run_{name}.npz"), status=res["status"], frames=0,
                                q_start=np.asarray(q0, dtype=float), **extra)
        print(name, "status", res["status"], "frames", len(res["snaps"]))

    # -- general SPD P0 / Q / R through the np.eye proxy ---------------------------------------
    P0, Q, R = spd(48, 1, 2.0), spd(48, 2, 0.5), spd(8, 3, 1.5)
    dense_frames = [0, 1, 2, 3, 4, 9, 19, 29, 39]
    for name, method, mparams in (("kf", "KF", dict(mp)), ("gmckf", "GMCKF", dict(mp, annealing=True)),
                                  ("imcckf", "IMCCKF", dict(mp, kernel_bw=30)),
                                  ("mckf", "MCKF", dict(mp, kernel_bw=50))):
        res = run_reference(ref_experiment, ref_noise, method, "GAUSSIAN_MIXTURE",
                            dict(std=1.0, mean=30.0, rho=0.1), q_a, mparams, t_max=2.0,
                            eye_queue=[P0, Q, R])
        extra = dict(method=method, noise_type="GAUSSIAN_MIXTURE", annealing=int(mparams["annealing"]),
                     kernel_bw0=float(mparams["kernel_bw"]), fpi_threshold=mparams["fpi_threshold"],
                     fpi_epoch_max=mparams["fpi_epoch_max"], hold_cnt=0, fov_check=0, P0=P0, Q=Q, R=R,
                     t_max=2.0)
        np.savez_compressed(os.path.join(OUT, f"ref_dense_{name}.npz"),
                            **pack_run(res, q_a, True, dense_frames, extra))
        print("dense", name, "status", res["status"], "frames", len(res["snaps"]))

    # -- closed-loop statistics of the reference itself (64 trials per estimator) --------------
    desired = np.array([149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0])
    for name, method, mparams in (("kf_alpha15", "KF", dict(mp)),
                                  ("gmckf_alpha15_anneal", "GMCKF", dict(mp, annealing=True))):
        n_stat = 256
        st = np.zeros((n_stat, 3)); status = np.zeros(n_stat, dtype=np.int64); q0s = np.zeros((n_stat, 6))
        for i in range(n_stat):
            q0 = noise_model.start_pose(q_nom, 12345, i, True)
            q0[1] = min(q0[1], -0.15)          # keep every start pose inside the field of view
            res = run_reference(ref_experiment, ref_noise, method, "ALPHA_STABLE",
                                dict(alpha=1.5, beta=0, gamma=1, delta=0), q0, mparams,
                                noise_seed=123456 + 977 * i)
            status[i] = res["status"]; q0s[i] = q0
            st[i] = stats_mod.trial_stats(res["t_log"], res["f_log"], desired)
        np.savez_compressed(os.path.join(OUT, f"ref_stats_{name}.npz"), stats=st, status=status, q0=q0s,
                            method=method, annealing=int(mparams["annealing"]))
        print("stats", name, "ITAE mean", st[:, 2].mean(), "fails", int(status.sum()))

    # -- noise distributions + hold trace -------------------------------------------------------
    NoiseType = ref_noise.NoiseType
    dists = {
        "white": ("WHITE_NOISE", dict(std=2.0)),
        "mixture": ("GAUSSIAN_MIXTURE", dict(std=1.0, mean=50.0, rho=0.1)),
        "bimodal": ("GAUSSIAN_BIMODAL", dict(std=1.0, mean=50.0, rho=0.2)),
        "alpha15": ("ALPHA_STABLE", dict(alpha=1.5, beta=0, gamma=1, delta=0)),
        "alpha10": ("ALPHA_STABLE", dict(alpha=1.0, beta=0, gamma=1, delta=0)),
        "alpha20": ("ALPHA_STABLE", dict(alpha=2.0, beta=0, gamma=1, delta=0)),
        "alpha15_skew": ("ALPHA_STABLE", dict(alpha=1.5, beta=0.5, gamma=2.0, delta=1.0)),
        "alpha10_skew": ("ALPHA_STABLE", dict(alpha=1.0, beta=0.5, gamma=2.0, delta=1.0)),
        "alpha05_levy": ("ALPHA_STABLE", dict(alpha=0.5, beta=1.0, gamma=1.0, delta=0.0)),
        "uniform": ("UNIFORM", dict()),
    }
    noise_out = {}
    for name, (ntype, nparams) in dists.items():
        prof = ref_noise.NoiseProfiler(8, NoiseType[ntype], seed=4242, noise_params=dict(nparams))
        draws = np.stack([prof.getNoise().copy() for _ in range(1250)]).ravel()
        noise_out["draws_" + name] = draws
        for k, v in nparams.items():
            noise_out[f"param_{name}_{k}"] = float(v)
    # hold trace: inject a known fresh stream into the reference profiler
    rng = np.random.default_rng(7)
    fresh = rng.standard_normal((400, 8)) * 12.0
    prof = ref_noise.NoiseProfiler(8, NoiseType.WHITE_NOISE, seed=1, noise_hold=True, noise_hold_cnt=10,
                                   noise_params=dict(std=1.0))
    # fresh[k, j] is what coordinate j would draw at frame k (counter-based indexing)
    clock = {"k": 0}
    prof.generators = [_InjectedGen(fresh[:, j], clock) for j in range(8)]
    held = np.zeros((300, 8))
    for k in range(300):
        clock["k"] = k
        held[k] = prof.getNoise()
    noise_out["hold_fresh"] = fresh
    noise_out["hold_out"] = held
    np.savez_compressed(os.path.join(OUT, "ref_noise.npz"), **noise_out)
    print("noise goldens written")


if __name__ == "__main__":
    main()
