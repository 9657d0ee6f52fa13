"""Drop-in for the reference's ``experiment.py`` (Method, ExperimentStatus, Experiment;
experiment.py:6-359) plus the batched ``BatchExperiment`` the B200 engine is built for.

``Experiment(...).run()`` keeps the reference signature and the 9-tuple it returns, but the
whole closed loop (noise, kinematic UR10 + pinhole camera, estimator update, IBVS law,
joint integration, error integrals) runs inside one CUDA kernel of libuvs_b200.so.
"""
import logging
from enum import Enum

import numpy as np

from . import _lib as L
from . import engine
from .ur10_simulation import camera_log_from_q


class Method(Enum):               # experiment.py:6-11
    ANALYTICAL = 1
    KF = 2
    MCKF = 3
    IMCCKF = 4
    GMCKF = 5
    RMCKF = 5                     # the paper's name for GMCKF (alias)


class ExperimentStatus(Enum):     # experiment.py:13-15
    SUCCESS = 0
    FAIL = 1


def _method_value(method):
    v = int(getattr(method, "value", method))
    if v == Method.ANALYTICAL.value:
        # experiment.py:121 dereferences an undefined R for ANALYTICAL (UnboundLocalError in the reference)
        raise NotImplementedError("Method.ANALYTICAL is not an estimator (and crashes in the reference)")
    if v not in (2, 3, 4, 5):
        raise ValueError(f"unknown method {method!r}")
    return v


class _Common:
    def _init_method(self, method, method_params):
        if "method_params" in method_params:          # experiment.py:29-30
            method_params = method_params["method_params"]
        self.method = method
        self._method = _method_value(method)
        self.initial_guess = method_params["initial_guess"]        # experiment.py:34 (KeyError if absent)
        self.kernel_bw, self.fpi_threshold, self.fpi_epoch_max, self.annealing = 10.0, 0.1, 1000, False
        if self._method in (3, 4, 5):                                # experiment.py:36-40
            self.kernel_bw = method_params["kernel_bw"]
            self.fpi_threshold = method_params["fpi_threshold"]
            self.fpi_epoch_max = method_params["fpi_epoch_max"]
            self.annealing = method_params["annealing"]
        # additive options of this engine
        self.mode = method_params.get("mode", L.MODE_AUTO)
        self.joseph = method_params.get("joseph", True)
        self.mckf_skip = method_params.get("mckf_skip", True)
        self.fov_check = method_params.get("fov_check", False)
        # additive: joint servo limits of the scenes (rad/s, rad/s^2; 0 = the kinematic plant), include/uvs_b200.h
        self.v_max = float(method_params.get("v_max", 0.0))
        self.a_max = float(method_params.get("a_max", 0.0))
        self.servo_gain = float(method_params.get("servo_gain", 0.0))
        # additive: the scene plant (sensor mount offset, finite discs, sub-stepped position servo), include/uvs_b200.h
        self.plant = dict(sensor_offset=method_params.get("sensor_offset"),
                          disc_radius=float(method_params.get("disc_radius", 0.0)),
                          servo_substeps=int(method_params.get("servo_substeps", 0)),
                          servo_kp=float(method_params.get("servo_kp", 0.0)))
        self.P0, self.Q, self.R = (method_params.get(k) for k in ("P0", "Q", "R"))

    def _filter_params(self, nf):
        return engine.filter_params(self._method, nf, mode=self.mode, joseph=self.joseph, mckf_skip=self.mckf_skip,
                                    kernel_bw=self.kernel_bw, fpi_threshold=self.fpi_threshold,
                                    fpi_epoch_max=self.fpi_epoch_max, annealing=self.annealing, Q=self.Q, R=self.R)

    def _setup_logger(self, logger):
        self.logger = logging.getLogger(__name__)
        if logger is not None:                         # experiment.py:42-46
            self.logger.setLevel(logger.level)
            for handler in logger.handlers:
                self.logger.addHandler(handler)


class Experiment(_Common):
    """experiment.py:17-359.  One closed-loop IBVS experiment."""

    def __init__(self, q_start: list, desired_f: list, noise_prof: object, t_s: float, t_max: float,
                 ibvs_gain: float, robot: object, method, logger: object = None, **method_params) -> None:
        self.q_start = q_start
        self.desired_f = desired_f
        self.robot = robot
        self.noise_prof = noise_prof
        self.t_s = t_s
        self.t_max = t_max
        self.ibvs_gain = ibvs_gain
        self._init_method(method, method_params)
        self._setup_logger(logger)

    def run(self) -> list:
        m = len(self.desired_f)
        nf = m // 2
        T = int(self.t_max / self.t_s)
        targets = getattr(self.robot, "targets", None)
        if self.robot is not None:
            self.robot.start(np.asarray(self.q_start, dtype=float))
        noise_in = None
        trial = 0
        if self.noise_prof is None:
            nparams = engine.noise_params(L.NOISE_NONE)
        elif hasattr(self.noise_prof, "params"):
            nparams = self.noise_prof.params()
            trial = self.noise_prof.trial
        else:
            # foreign profiler (e.g. the reference's PCG64 one): draw its T vectors and replay them
            nparams = engine.noise_params(L.NOISE_NONE)
            noise_in = engine.dev(np.stack([np.array(self.noise_prof.getNoise(), dtype=float) for _ in range(T)])[None])
        cfg = engine.closed_loop_cfg(self._filter_params(nf), nparams, T=T, dt=self.t_s, ibvs_gain=self.ibvs_gain,
                                     q_start=self.q_start, desired_f=self.desired_f, targets=targets,
                                     initial_guess=self.initial_guess, change_q_start=False,
                                     fov_check=self.fov_check, P0=self.P0, v_max=self.v_max, a_max=self.a_max, servo_gain=self.servo_gain, **self.plant)
        out = engine.closed_loop(cfg, trial, 1, want_log=True, noise_in=noise_in,
                                 q0_in=engine.dev(np.asarray(self.q_start, dtype=float).reshape(1, 6)))
        k = int(out["frames"][0].item())
        log = out["log"][0, :k].cpu().numpy()
        status = ExperimentStatus(int(out["status"][0].item()))
        q_log, f_log, noise_log = log[:, 0:6], log[:, 6:6 + m], log[:, 6 + m:6 + 2 * m]
        desired = np.asarray(self.desired_f, dtype=float)
        t_log = (np.arange(k) + 1) * self.t_s
        desired_f_log = np.tile(desired, (k, 1))
        error_log = f_log - desired[None, :]                          # experiment.py:302,327
        camera_log = camera_log_from_q(q_log)                          # experiment.py:324
        if self._method == Method.MCKF.value:                          # experiment.py:330
            kk = np.arange(k)
            kernel_bw_log = self.kernel_bw + (100 * (1 - kk / T) if self.annealing else 0.0 * kk)
        else:
            kernel_bw_log = -np.ones(k)
        if self.robot is not None and k > 0 and hasattr(self.robot, "setJointsPos"):
            self.robot.setJointsPos(q_log[-1])
            self.robot.step()
            self.robot.stop()
        if status == ExperimentStatus.SUCCESS:
            self.logger.info("Experiment success")
        else:
            self.logger.error("Experiment failed")
        return status, t_log, error_log, q_log, f_log, desired_f_log, camera_log, noise_log, kernel_bw_log


class BatchExperiment(_Common):
    """B independent closed-loop experiments in one launch (additive API).

    Trial b of the batch is global trial ``trial0 + b``: its start pose is main.py:129-134 drawn from the Philox
    stream keyed by ``pose_seed``, its noise the stream keyed by ``noise_prof.seed``; with ``sweep`` the noise
    level of a trial is ``sweep[trial // epoch]`` (main.py:121-127).
    """

    def __init__(self, q_start, desired_f, noise_prof, t_s, t_max, ibvs_gain, method, *, n_trials, trial0=0,
                 change_q_start=True, pose_seed=12345, sweep=None, sweep_epoch=1, sweep_param=None, targets=None,
                 logger=None, **method_params):
        self.q_start, self.desired_f, self.noise_prof = q_start, desired_f, noise_prof
        self.t_s, self.t_max, self.ibvs_gain = t_s, t_max, ibvs_gain
        self.n_trials, self.trial0 = int(n_trials), int(trial0)
        self.change_q_start, self.pose_seed = change_q_start, pose_seed
        self.sweep, self.sweep_epoch, self.targets = sweep, sweep_epoch, targets
        if sweep is not None and sweep_param is None:
            sweep_param = "alpha" if noise_prof is not None and noise_prof.noise_type.name == "ALPHA_STABLE" else "rho"
        self.sweep_param = {None: 0, "rho": 1, "alpha": 2}[sweep_param]
        self._init_method(method, method_params)
        self._setup_logger(logger)

    def config(self):
        nf = len(self.desired_f) // 2
        nparams = self.noise_prof.params() if self.noise_prof is not None else engine.noise_params(L.NOISE_NONE)
        return engine.closed_loop_cfg(self._filter_params(nf), nparams, T=int(self.t_max / self.t_s), dt=self.t_s,
                                      ibvs_gain=self.ibvs_gain, q_start=self.q_start, desired_f=self.desired_f,
                                      targets=self.targets, initial_guess=self.initial_guess,
                                      change_q_start=self.change_q_start, fov_check=self.fov_check,
                                      pose_seed=self.pose_seed, sweep=self.sweep, sweep_epoch=self.sweep_epoch,
                                      sweep_param=self.sweep_param, P0=self.P0, v_max=self.v_max, a_max=self.a_max, servo_gain=self.servo_gain, **self.plant)

    def run(self, want_log=False, want_state=False, to_host=True, **kw):
        out = engine.closed_loop(self.config(), self.trial0, self.n_trials, want_log=want_log,
                                 want_state=want_state, **kw)
        if to_host:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        return out
