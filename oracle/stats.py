"""Oracle (test infrastructure): per-trial error statistics.

Follows /root/reference/results/plot_errorbar.m:39-94: for each experiment,
e_j = desired_f_j - f_j (the LOGGED, noisy feature), then
    ISE  = || (sum_k e_jk^2)_j ||_2
    IAE  = || (sum_k |e_jk|)_j ||_2
    ITAE = || (sum_k t_k |e_jk|)_j ||_2
and per noise level mean / median / std (MATLAB std = N-1) over SUCCESS runs.
"""
import numpy as np


def trial_stats(t_log, f_log, desired_f):
    e = np.asarray(desired_f)[None, :] - np.asarray(f_log)
    ise = np.linalg.norm(np.sum(e * e, axis=0))
    iae = np.linalg.norm(np.sum(np.abs(e), axis=0))
    itae = np.linalg.norm(np.asarray(t_log) @ np.abs(e))
    return np.array([ise, iae, itae])


def config_stats(stats, status):
    """stats [B,3], status [B] (0 = SUCCESS) -> dict of mean/median/std/n."""
    ok = np.asarray(status) == 0
    s = np.asarray(stats)[ok]
    n = int(ok.sum())
    out = {"n_ok": n, "n_fail": int((~ok).sum())}
    for j, name in enumerate(("ise", "iae", "itae")):
        col = s[:, j] if n else np.zeros(0)
        out[name + "_mean"] = float(np.mean(col)) if n else float("nan")
        out[name + "_median"] = float(np.median(col)) if n else float("nan")
        out[name + "_std"] = float(np.std(col, ddof=1)) if n > 1 else float("nan")
    return out
