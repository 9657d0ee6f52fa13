"""Oracle (test infrastructure): golden vectors of the camera-twist pipeline, minted by EXECUTING the
reference's own tests/ scripts, unmodified, against a stubbed simulator.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden_twist          # writes tests/golden/ref_twist_<script>.npz

Each /root/reference/tests/<script>.py is a module-level program that talks to CoppeliaSim, waits on
``input()``, plots with matplotlib (absent here) and writes a CSV.  It is compiled from the file where
it lies and executed in a fresh namespace with these stand-ins installed beforehand:

  ur10_simulation.UR10Simulation  -> ``ScriptRobot`` (kinematic plant of oracle/fake_robot.py; the scene
                                     is chosen per script: 1, 3 or 4 discs)
  utils.detect{GreenCircle,RGBCircles,4Circles} -> identity on the projected feature vector
  matplotlib / matplotlib.pyplot / zmqRemoteApi -> inert stubs;  builtins.input -> "" ;  DataFrame.to_csv -> no-op
  noise.NoiseProfiler             -> ``ScriptNoise``: the *_4_features scripts call
                                     ``NoiseProfiler(m, GAUSSIAN_MIXTURE, rho=..)`` and
                                     ``.getBimodalGaussianMixture()``, which the shipped noise.py no
                                     longer accepts (KeyError 'std' at noise.py:46; the method is private
                                     at :140).  The stand-in draws the bimodal mixture of noise.py:140-150
                                     with std = 1, mean = 50 from a seeded NumPy generator; the stream is
                                     what the script itself logs in ``noise_log`` and is replayed verbatim
                                     by the oracle and the CUDA kernel.
  numpy.random.seed(SEED)         before the script runs (X = np.random.rand(..) at kalman_3_features.py:50)

Nothing of the scripts is copied: only the arrays they compute (q_log, f_log, X_log, final X / P, the
noise they consumed) are stored.  The simulated clock starts at 3 s (so the scripts' "wait for the robot"
loops fall through) and jumps past T_MAX after ``max_frames`` frames to bound the file sizes.
"""
import builtins
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 20240607

from . import filters, kinematics                                  # noqa: E402
from .fake_robot import FakeRobot                                   # noqa: E402
from . import twist_loop as TL                                      # noqa: E402

# script -> (nf, max_frames, oracle configuration that restates the script's constants)
SCRIPTS = {
    "ibvs_1_feature": dict(nf=1, max_frames=240, cfg=dict(
        method=filters.ANALYTICAL, gain=0.75, desired_f=[128.0, 128.0], focal=512.0, depth_mode=TL.DEPTH_HEIGHT,
        x0_mode=TL.X0_ZEROS)),
    "ibvs_3_features": dict(nf=3, max_frames=240, cfg=dict(
        method=filters.ANALYTICAL, gain=0.75, focal=1024.0, x0_mode=TL.X0_ZEROS)),
    "ibvs_4_features": dict(nf=4, max_frames=440, cfg=dict(
        method=filters.ANALYTICAL, gain=0.5, focal=256.0 / (2.0 * np.tan(0.5 * 0.65)), x0_mode=TL.X0_ZEROS)),
    "kalman_1_feature": dict(nf=1, max_frames=600, cfg=dict(
        method=filters.KF, gain=0.05, desired_f=[128.0, 128.0], x0_mode=TL.X0_ZEROS, joseph=False)),
    "kalman_3_features": dict(nf=3, max_frames=600, cfg=dict(
        method=filters.KF, gain=0.05, x0_mode=TL.X0_RANDOM, joseph=False, err_threshold=2.0,
        dp_zero_mask=0b011100, square_inv=True)),
    "kalman_4_features": dict(nf=4, max_frames=440, cfg=dict(
        method=filters.KF, gain=0.5, x0_mode=TL.X0_IMAGE_JACOBIAN, focal_x0=1024.0, joseph=True,
        err_threshold=0.01, stale_fk=True)),
    "mckf_1_feature": dict(nf=1, max_frames=400, cfg=dict(
        method=filters.MCKF, gain=0.001, desired_f=[128.0, 128.0], x0_mode=TL.X0_RANDOM, joseph=True,
        err_threshold=0.01, kernel_bw=200.0, fpi_threshold=0.01, fpi_epoch_max=100, mckf_skip=False, stale_fk=True)),
    "mckf_3_features": dict(nf=3, max_frames=300, cfg=dict(
        method=filters.MCKF, gain=0.001, x0_mode=TL.X0_RANDOM, joseph=True, square_inv=True,
        err_threshold=0.01, kernel_bw=200.0, fpi_threshold=0.01, fpi_epoch_max=100, mckf_skip=False, stale_fk=True)),
    "mckf_4_features": dict(nf=4, max_frames=300, cfg=dict(
        method=filters.MCKF, gain=0.5, x0_mode=TL.X0_IMAGE_JACOBIAN,
        focal_x0=256.0 / (2.0 * np.tan(0.5 * 0.65)), joseph=True, err_threshold=0.01, kernel_bw=20.0,
        fpi_threshold=0.01, fpi_epoch_max=100, mckf_skip=False, stale_fk=True)),
}


class ScriptRobot(FakeRobot):
    """FakeRobot with the scripts' clock: t = 3 s + frames * dt, then +inf after max_frames."""
    NF = 4
    MAX_FRAMES = 400

    def __init__(self):
        super().__init__(dt=0.05, t_max=1e9, nf=self.NF, snapshot=False)
        self.q_cmd_log = []

    def _time(self):
        frames = self.steps - 1
        return 3.0 + frames * self.dt if frames < self.MAX_FRAMES else float("inf")

    def setJointsPos(self, q):
        self.q_cmd = np.array(q, dtype=float)


class ScriptNoise:
    """Stand-in for the NoiseProfiler interface the *_4_features scripts were written against."""
    STD, MEAN = 1.0, 50.0

    def __init__(self, num_features, noise_type, rho=0.0, **kw):
        self.m = num_features
        self.rho = rho
        self.rng = np.random.Generator(np.random.PCG64(SEED))

    def getBimodalGaussianMixture(self):                    # noise.py:140-150, per coordinate
        out = np.zeros(self.m)
        for i in range(self.m):
            u = self.rng.uniform()
            loc = 0.0 if u > self.rho else (self.MEAN if u > self.rho / 2 else -self.MEAN)
            out[i] = self.rng.normal(loc=loc, scale=self.STD)
        return out


class _Inert(types.ModuleType):
    """matplotlib stand-in: every attribute is a callable that returns another inert object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert(name)

    def __call__(self, *a, **kw):
        return _Inert("call")

    def __iter__(self):
        return iter(())


def run_script(name, nf, max_frames, rho=None):
    if not os.path.isdir(REF):
        raise SystemExit("reference tree not present; goldens can only be minted in the build container")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import pandas as pd
    import noise as ref_noise
    import ur10_simulation as ref_sim
    import utils as ref_utils

    ScriptRobot.NF, ScriptRobot.MAX_FRAMES = nf, max_frames
    robots = []

    def make_robot(*a, **kw):
        robots.append(ScriptRobot())
        return robots[-1]

    ident = lambda image: np.array(image, dtype=float).copy()       # noqa: E731
    saved = dict(sim=ref_sim.UR10Simulation, prof=ref_noise.NoiseProfiler, input=builtins.input,
                 to_csv=pd.DataFrame.to_csv, mpl=sys.modules.get("matplotlib"), plt=sys.modules.get("matplotlib.pyplot"),
                 det=[getattr(ref_utils, n) for n in ("detectGreenCircle", "detectRGBCircles", "detect4Circles")])
    ref_sim.UR10Simulation = make_robot
    ref_noise.NoiseProfiler = ScriptNoise
    for n in ("detectGreenCircle", "detectRGBCircles", "detect4Circles"):
        setattr(ref_utils, n, ident)
    builtins.input = lambda *a: ""
    pd.DataFrame.to_csv = lambda self, *a, **kw: None
    mpl = _Inert("matplotlib")
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = _Inert("matplotlib.pyplot")
    mpl.pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["zmqRemoteApi"] = _Inert("zmqRemoteApi")      # legacy client name imported (unused) by ibvs_1_feature.py:9
    path = os.path.join(REF, "tests", name + ".py")
    ns = {"__name__": "__main__", "__file__": path}
    np.random.seed(SEED)
    x0 = None
    try:
        with open(path) as fh:
            code = compile(fh.read(), path, "exec")
        real_print = builtins.print
        builtins.print = lambda *a, **kw: None
        # X0: the scripts that start from np.random.rand draw it right after the seed; re-derive it the same way
        try:
            exec(code, ns)
        finally:
            builtins.print = real_print
    finally:
        ref_sim.UR10Simulation = saved["sim"]
        ref_noise.NoiseProfiler = saved["prof"]
        builtins.input = saved["input"]
        pd.DataFrame.to_csv = saved["to_csv"]
        for n, fn in zip(("detectGreenCircle", "detectRGBCircles", "detect4Circles"), saved["det"]):
            setattr(ref_utils, n, fn)
        sys.modules.pop("zmqRemoteApi", None)
        for key, mod in (("matplotlib", saved["mpl"]), ("matplotlib.pyplot", saved["plt"])):
            if mod is None:
                sys.modules.pop(key, None)
            else:
                sys.modules[key] = mod
    k = int(ns["k"])
    m = 2 * nf
    out = dict(frames=k, q_log=np.array(ns["q_log"])[:k], f_log=np.array(ns["f_log"])[:k],
               q_final=np.array(robots[0].q_cmd), dt=float(ns["TS"]), gain=float(ns["GAIN"]))
    out["noise_log"] = np.array(ns["noise_log"])[:k] if "noise_log" in ns else np.zeros((k, m))
    if "X" in ns and "P" in ns:
        out["X_final"] = np.array(ns["X"], dtype=float).ravel()
        out["P_final"] = np.array(ns["P"], dtype=float)
    if "X_log" in ns:
        out["X_log"] = np.array(ns["X_log"])[:k]
    return out, x0


def script_x0(cfg, N):
    """The scripts' X = np.random.rand(N, 1) is the first draw after np.random.seed(SEED)."""
    if cfg.get("x0_mode") != TL.X0_RANDOM:
        return None
    np.random.seed(SEED)
    return np.random.rand(N, 1).ravel()


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, spec in SCRIPTS.items():
        nf = spec["nf"]
        gold, _ = run_script(name, nf, spec["max_frames"])
        x0 = script_x0(spec["cfg"], 12 * nf)
        cfg = TL.make_cfg(nf=nf, T=spec["max_frames"], **spec["cfg"])
        if nf != 4:
            cfg["desired_f"] = spec["cfg"].get("desired_f", [153.5, 144.5, 125.5, 116.5, 99.5, 144.5])
        # the oracle restatement on the same noise stream (padded to T frames)
        noise = np.zeros((spec["max_frames"], 2 * nf))
        noise[:gold["frames"]] = gold["noise_log"]
        res = TL.run_trial(cfg, 0, noise=noise, x0=x0, record=True)
        dq = np.abs(res["q_log"][:gold["frames"]] - gold["q_log"]).max() if res["frames"] >= gold["frames"] else np.inf
        print(f"{name:20s} frames ref {gold['frames']:4d} oracle {res['frames']:4d}  max|dq_log| {dq:.2e}")
        save = {k: v for k, v in gold.items()}
        save["cfg_keys"] = np.array(sorted(spec["cfg"].keys()))
        for k2, v in spec["cfg"].items():
            save["cfg_" + k2] = np.asarray(v, dtype=float)
        save["nf"] = nf
        save["T"] = spec["max_frames"]
        save["desired_f"] = np.asarray(cfg["desired_f"], dtype=float)[:2 * nf]
        save["q_start"] = np.asarray(cfg["q_start"], dtype=float)
        if x0 is not None:
            save["x0"] = x0
        np.savez_compressed(os.path.join(OUT, f"ref_twist_{name}.npz"), **save)


if __name__ == "__main__":
    main()
