"""uvs_b200 — B200-native Monte Carlo engine for the uncalibrated-visual-servoing hot path.

The directory name carries a hyphen (it mirrors the reference repository's name), so import it
with ``importlib.import_module("uncalibrated-visual-servoing_b200")`` or through the ``uvs_b200``
alias module at the repository root.
"""
from . import _lib, engine, sweep                               # noqa: F401
from .experiment import BatchExperiment, Experiment, ExperimentStatus, Method   # noqa: F401
from .noise import NoiseProfiler, NoiseType                    # noqa: F401
from .ur10_simulation import UR10Simulation                    # noqa: F401

__all__ = ["engine", "sweep", "Experiment", "BatchExperiment", "ExperimentStatus", "Method", "NoiseProfiler", "NoiseType",
           "UR10Simulation"]
