"""Oracle (test infrastructure): duck-typed stand-ins for the reference's
``UR10Simulation`` (/root/reference/ur10_simulation.py:7-211) so that the
UNMODIFIED ``experiment.Experiment.run()`` can be executed without CoppeliaSim
(SURVEY.md §8c, Appendix C).

``FakeRobot``   kinematic plant: q <- commanded q on ``step()``, features from
                the pinhole projection of the disc centres; ``getCameraImage``
                returns the feature vector itself (the caller patches
                ``experiment.detect4Circles`` to the identity).
A snapshot hook in ``setJointsPos`` (called once per frame from
experiment.py:335) copies the filter's frame locals X, P, K, dq, ... out of the
reference's own stack frame.
"""
import sys

import numpy as np

from . import kinematics


class _Sim:
    def __init__(self, robot):
        self._robot = robot

    def getSimulationTime(self):
        return self._robot._time()


class FakeRobot:
    perspective_angle = kinematics.PERSPECTIVE_ANGLE_DEG      # ur10_simulation.py:23

    def __init__(self, dt=0.05, t_max=15.0, nf=4, targets=None, fov_check=False, snapshot=True):
        self.dt = dt
        self.t_max = t_max
        self.n_frames = int(t_max / dt)
        self.targets = np.asarray(targets) if targets is not None else kinematics.TARGETS[nf]
        self.fov_check = fov_check
        self.sim = _Sim(self)
        self.q = np.zeros(6)
        self.q_cmd = np.zeros(6)
        self.steps = 0
        self._Ts = None                       # cached T_0_1 .. T_0_6 (ur10_simulation.py:59,97-110)
        self.snapshots = [] if snapshot else None

    # -- clock: frame k is seen at t_k = (k+1)*dt, exactly n_frames frames ---------------------
    def _time(self):
        k = self.steps - 1
        if k >= self.n_frames:
            return float(self.t_max)
        t = (k + 1) * self.dt
        return t if t < self.t_max else float(np.nextafter(self.t_max, 0.0))

    # -- ur10_simulation.py:35-66 ----------------------------------------------------------------
    def start(self, q=None):
        if q is not None:
            self.q = np.array(q, dtype=float)
        self.q_cmd = self.q.copy()
        self.steps = 0
        self.step()
        self.fkine(recalculate=True)          # ur10_simulation.py:59

    def stop(self):
        pass

    def step(self):
        self.q = self.q_cmd.copy()
        self.steps += 1

    def getJointsPos(self):
        return self.q

    def setJointsPos(self, q):
        self.q_cmd = np.array(q, dtype=float)
        if self.snapshots is not None:
            fr = sys._getframe(1)
            if fr.f_code.co_name == "run":
                L = fr.f_locals
                snap = {}
                for name in ("X", "P", "K", "H", "Z", "dq", "f", "k", "kernel_bw", "epoch",
                             "skip_correction", "kappa", "t"):
                    if name in L:
                        v = L[name]
                        snap[name] = np.array(v, dtype=float).copy() if isinstance(v, np.ndarray) else v
                self.snapshots.append(snap)

    # -- ur10_simulation.py:97-195 ------------------------------------------------------------
    # UR10Simulation caches T_0_1 .. T_0_6 and refreshes them ONLY in fkine(recalculate=True) (which first reads the
    # joints, :98-99); jacobian / getCameraRotation / getCameraPosition / getCameraHeight / computeZ pass their
    # recalculate_fkine flag on and otherwise work on the cache; computePose never touches it (it asks the simulator,
    # :156-168).  start() ends with fkine(recalculate=True) (:59).  Scripts that call robot.jacobian() /
    # getCameraRotation() without the flag therefore use the pose of the last refresh, not the current one.
    def fkine(self, recalculate=False, all_transforms=False):
        if recalculate or self._Ts is None:
            self._Ts = kinematics.fkine_all(self.q)
        return tuple(reversed(self._Ts)) if all_transforms else self._Ts[-1]

    def jacobian(self, recalculate_fkine=False):
        self.fkine(recalculate=recalculate_fkine)
        return kinematics.jacobian_from_transforms(self._Ts)

    def getCameraRotation(self, recalculate_fkine=False):
        return self.fkine(recalculate=recalculate_fkine)[0:3, 0:3]

    def getCameraPosition(self, recalculate_fkine=False):
        return self.fkine(recalculate=recalculate_fkine)[0:3, 3]

    def getCameraHeight(self, recalculate_fkine=False):
        return self.fkine(recalculate=recalculate_fkine)[2, 3]

    def computePose(self, recalculate_fkine=False):
        return kinematics.camera_pose(self.q)            # simulator query: always the current pose

    def computeZ(self, n=1, recalculate_fkine=False):
        cam = self.getCameraPosition(recalculate_fkine)
        return np.array([np.linalg.norm(cam - pw) for pw in self.targets[:n]])

    def getCameraImage(self):
        f = kinematics.project(self.q, self.targets)
        if self.fov_check:
            z = np.repeat(kinematics.depth(self.q, self.targets), 2)
            bad = ~((f >= 0.0) & (f < kinematics.RESOLUTION) & (z > 0.01))
            pair_bad = bad.reshape(-1, 2).any(axis=1).repeat(2)
            f = np.where(pair_bad, np.nan, f)     # empty colour mask -> 0/0 centroid (utils.py:73-78)
        return f, [kinematics.RESOLUTION, kinematics.RESOLUTION]


class StreamNoise:
    """Stands in for noise.NoiseProfiler: serves a precomputed [T,m] stream
    through getNoise() (noise.py:81-118 returns one vector per frame)."""

    def __init__(self, stream):
        self.stream = np.asarray(stream, dtype=float)
        self.k = 0

    def getNoise(self):
        v = self.stream[min(self.k, len(self.stream) - 1)].copy()
        self.k += 1
        return v
