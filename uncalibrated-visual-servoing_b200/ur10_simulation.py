"""Drop-in for the reference's ``ur10_simulation.UR10Simulation`` with CoppeliaSim
replaced by the kinematic model that the CUDA kernels integrate
(ur10_simulation.py:7-211; scene constants SURVEY.md Appendix A).

Joint commands take effect on ``step()`` (q <- commanded q), the camera frame is
DH frame 6, and ``getCameraImage()`` returns the projected feature vector in
place of an image (this engine has no renderer; ``detect4Circles`` is the
identity here).  Kinematics are evaluated by ``uvs_ur10_fk`` / ``uvs_project``.
"""
import logging

import numpy as np
import torch

from . import engine


class _Sim:
    def __init__(self, robot):
        self._robot = robot

    def getSimulationTime(self):
        return self._robot.steps * self._robot.dt


class UR10Simulation:
    perspective_angle = 65           # ur10_simulation.py:23
    resolution = [256, 256]

    def __init__(self, logger: object = None, visualization: bool = True, nf: int = 4, targets=None,
                 dt: float = 0.05) -> None:
        self.nf = int(nf)
        self.targets = np.asarray(engine.TARGETS[self.nf] if targets is None else targets, dtype=float)
        self.dt = float(dt)
        self.q = np.zeros(6)
        self.dq = np.zeros(6)
        self._q_cmd = np.zeros(6)
        self._T06 = self._Jg = self._q_fk = None      # cached transforms (ur10_simulation.py:59,97-110)
        self.steps = 0
        self.visualization = visualization
        self.sim = _Sim(self)
        self.logger = logging.getLogger(__name__)
        if logger is not None:
            self.logger.setLevel(logger.level)
            for handler in logger.handlers:
                self.logger.addHandler(handler)

    # -- simulation control (ur10_simulation.py:35-66,201-202) -------------------------------
    def start(self, q: list = None):
        if q is not None:
            self.q = np.array(q, dtype=float)
        self._q_cmd = self.q.copy()
        self.steps = 0
        self.step()
        self.fkine(recalculate=True, all_transforms=True)          # ur10_simulation.py:59

    def stop(self):
        pass

    def step(self):
        self.dq = (self._q_cmd - self.q) / self.dt
        self.q = self._q_cmd.copy()
        self.steps += 1

    def setJointsPos(self, q):
        self._q_cmd = np.array(q, dtype=float)

    def getJointsPos(self):
        return self.q

    def getJointsVel(self):
        return self.dq

    # -- analytic kinematics (ur10_simulation.py:97-195) -----------------------------------------
    # Like the reference, the transforms are CACHED: fkine(recalculate=True) reads the joints and refreshes them
    # (ur10_simulation.py:97-105), start() ends with such a refresh (:59), and every other accessor works on the cache
    # unless asked to recalculate — a script that calls jacobian() / getCameraRotation() without the flag gets the pose
    # of the last refresh, exactly as with the reference object.  T_0_6 and the geometric Jacobian come from
    # uvs_ur10_fk; the intermediate T_0_1..T_0_5 (all_transforms=True, used by nothing on the hot path) are the same
    # DH products evaluated on the host.
    _DH = ((0.0, 0.128, 0.0, -np.pi / 2), (-np.pi / 2, 0.0, 0.6127, 0.0), (0.0, 0.0, 0.5716, 0.0),
           (-np.pi / 2, 0.1639, 0.0, -np.pi / 2), (0.0, 0.1157, 0.0, np.pi / 2), (np.pi, 0.0922, 0.0, 0.0))

    @staticmethod
    def dh(theta, d, a, alpha):
        """ur10_simulation.py:204-211: Rz(theta) Tz(d) Tx(a) Rx(alpha)."""
        ct, st, ca, sa = np.cos(theta), np.sin(theta), np.cos(alpha), np.sin(alpha)
        return np.array([[ct, -st * ca, st * sa, a * ct], [st, ct * ca, -ct * sa, a * st],
                         [0.0, sa, ca, d], [0.0, 0.0, 0.0, 1.0]])

    def _refresh(self):
        T06, Jg = engine.ur10_fk(engine.dev(self.q.reshape(1, 6)))
        self._T06, self._Jg = T06[0].cpu().numpy(), Jg[0].cpu().numpy()
        self._q_fk = self.q.copy()

    def _cached(self, recalculate):
        if recalculate or self._T06 is None:
            self._refresh()
        return self._T06, self._Jg

    def fkine(self, recalculate=False, all_transforms=False):
        T06, _ = self._cached(recalculate)
        if not all_transforms:
            return T06
        Ts, T = [], np.eye(4)
        for i, (off, d, a, alpha) in enumerate(self._DH):          # ur10_simulation.py:100-105
            T = T @ self.dh(self._q_fk[i] + off, d, a, alpha)
            Ts.append(T)
        Ts[5] = T06
        return tuple(reversed(Ts))                                  # T_0_6, T_0_5, ..., T_0_1 (:107-108)

    def jacobian(self, recalculate_fkine=False):
        return self._cached(recalculate_fkine)[1]

    def getCameraRotation(self, recalculate_fkine=False):
        return self._cached(recalculate_fkine)[0][0:3, 0:3]

    def getCameraHeight(self, recalculate_fkine=False):
        return self._cached(recalculate_fkine)[0][2, 3]

    def getCameraPosition(self, recalculate_fkine=False):
        return self._cached(recalculate_fkine)[0][0:3, 3]

    def computePose(self, recalculate_fkine=False):
        # the reference asks the simulator for the sensor pose (ur10_simulation.py:156-168): always the current joints
        T06, _ = engine.ur10_fk(engine.dev(self.q.reshape(1, 6)), want_jacobian=False)
        return camera_pose_from_T(T06.cpu().numpy())[0]

    def computeZ(self, n=1, recalculate_fkine=False):
        t = self.getCameraPosition(recalculate_fkine)
        return np.array([np.linalg.norm(t - self.targets[i]) for i in range(n)])

    def getCameraImage(self):
        f, _ = engine.project(self.nf, engine.dev(self.q.reshape(1, 6)), self.targets)
        return f[0].cpu().numpy(), list(self.resolution)


def camera_pose_from_T(T06):
    """[B,4,4] -> [B,6] sensor position + roll/pitch/yaw.  The vision sensor is DH-6 . Rz(pi);
    the angles are utils.py:174-179 applied to its quaternion (ur10_simulation.py:156-168)."""
    T06 = np.asarray(T06, dtype=float)
    Rs = T06[:, 0:3, 0:3] * np.array([-1.0, -1.0, 1.0])[None, None, :]
    pose = np.zeros((T06.shape[0], 6))
    pose[:, 0:3] = T06[:, 0:3, 3]
    pose[:, 3] = np.arctan2(Rs[:, 2, 1], Rs[:, 2, 2])
    pose[:, 4] = np.arcsin(np.clip(-Rs[:, 2, 0], -1.0, 1.0))
    pose[:, 5] = np.arctan2(Rs[:, 1, 0], Rs[:, 0, 0])
    return pose


def camera_log_from_q(q_log):
    """camera_log of experiment.py:324 for a [k,6] joint log."""
    q = np.asarray(q_log, dtype=float).reshape(-1, 6)
    if len(q) == 0:
        return np.zeros((0, 6))
    T06, _ = engine.ur10_fk(engine.dev(q), want_jacobian=False)
    return camera_pose_from_T(T06.cpu().numpy())
