"""Drop-in for the reference's ``noise.py`` (NoiseType, NoiseProfiler; noise.py:7-207).

Same constructor and ``getNoise()`` contract, but the draws come from the
counter-based Philox4x32-10 streams of libuvs_b200 (``uvs_noise_fill``): sample
k of coordinate j is a pure function of (seed, trial, k, j), so the same
profiler object can also be handed to the batched GPU engine, which then
generates exactly the same stream inside the closed-loop kernel.
"""
import logging
import os
from enum import Enum

import numpy as np

from . import _lib as L
from . import engine


class NoiseType(Enum):            # noise.py:7-12
    WHITE_NOISE = 1
    GAUSSIAN_MIXTURE = 2
    GAUSSIAN_BIMODAL = 3
    ALPHA_STABLE = 4
    UNIFORM = 5

    @classmethod
    def numberOfGenerators(cls, type) -> int:     # noise.py:14-27 (kept for API compatibility)
        return {cls.WHITE_NOISE: 1, cls.GAUSSIAN_MIXTURE: 2, cls.GAUSSIAN_BIMODAL: 3, cls.ALPHA_STABLE: 1,
                cls.UNIFORM: 1}.get(type, 0)


class NoiseProfiler:
    """noise.py:29-118.  ``trial`` selects the Philox sub-stream (default 0)."""

    def __init__(self, num_features: int, noise_type, seed: int = None, logger: object = None,
                 noise_hold: bool = False, noise_hold_cnt: int = 0, trial: int = 0, **noise_params) -> None:
        self.num_features = int(num_features)
        self.logger = logging.getLogger(__name__)
        if logger is not None:
            self.logger.setLevel(logger.level)
            for handler in logger.handlers:
                self.logger.addHandler(handler)
        if not isinstance(noise_type, NoiseType):
            noise_type = NoiseType(int(getattr(noise_type, "value", noise_type)))
        self.noise_type = noise_type
        if "noise_params" in noise_params:           # noise.py:43-44
            noise_params = noise_params["noise_params"]
        self.std = self.mean = self.rho = 0.0
        self.alpha, self.beta, self.gamma, self.delta = 2.0, 0.0, 1.0, 0.0
        if noise_type in (NoiseType.WHITE_NOISE, NoiseType.GAUSSIAN_MIXTURE, NoiseType.GAUSSIAN_BIMODAL):
            self.std = noise_params["std"]           # KeyError like noise.py:47
            if noise_type != NoiseType.WHITE_NOISE:
                self.mean = noise_params["mean"]
                self.rho = noise_params["rho"]
        elif noise_type == NoiseType.ALPHA_STABLE:
            self.alpha = noise_params["alpha"]
            self.beta = noise_params["beta"]
            self.gamma = noise_params["gamma"]
            self.delta = noise_params["delta"]
        self.seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self.trial = int(trial)
        self.noise_hold_cnt_max_predefined = int(noise_hold_cnt) if noise_hold else 0   # noise.py:72-75
        self.noise = np.zeros(self.num_features)
        self._k = 0
        self._chunk = None

    def params(self):
        """The C-ABI noise block for this profiler."""
        return engine.noise_params(self.noise_type.value, seed=self.seed, hold_cnt=self.noise_hold_cnt_max_predefined,
                                   std=self.std, mean=self.mean, rho=self.rho, alpha=self.alpha, beta=self.beta,
                                   gamma=self.gamma, delta=self.delta)

    def stream(self, T):
        """First T frames of this profiler's stream as a [T, num_features] numpy array."""
        out = engine.noise_fill(self.params(), self.trial, 1, int(T), self.num_features)
        return out[0].cpu().numpy()

    def getNoise(self) -> np.ndarray:
        """noise.py:81-118 — one vector per call; returns the SAME array object each call."""
        if self._chunk is None or self._k >= len(self._chunk):
            self._chunk = self.stream(max(1024, 2 * (self._k + 1)))
        self.noise[:] = self._chunk[self._k]
        self._k += 1
        return self.noise
