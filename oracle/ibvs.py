"""Oracle (test infrastructure): the IBVS control step.

Follows /root/reference/experiment.py:300-312:
    J = X.reshape(m, 6); error = f - desired_f
    dq = -gain * pinv(J) @ (kappa * error)
np.linalg.pinv = SVD with singular values <= 1e-15*s_max zeroed.
"""
import numpy as np


def ibvs_step(X, f, desired_f, kappa, gain):
    m = len(f)
    J = np.asarray(X, dtype=float).reshape(m, 6)
    error = np.asarray(f, dtype=float) - np.asarray(desired_f, dtype=float)
    b = (np.asarray(kappa, dtype=float) * error).reshape(m, 1)
    dq = -gain * np.linalg.pinv(J) @ b
    return dq.ravel(), error
