"""Oracle (test infrastructure): feature-noise models on counter-based Philox.

Distributions follow /root/reference/noise.py:120-207 (white, Gaussian mixture,
bimodal mixture, alpha-stable by Chambers-Mallows-Stuck, uniform) and the
per-feature-pair outlier hold of noise.py:81-118.  The reference draws from
numpy PCG64 generators whose streams overlap across experiments
(SURVEY.md §3.3); here every draw is a pure function of
(seed, trial, frame, coordinate, draw) through Philox4x32-10, which is what the
CUDA kernels implement too, so GPU-vs-oracle noise is comparable sample by
sample while reference-vs-oracle is comparable in distribution only.

Counter layout (shared with csrc/philox.cuh):
    key = (seed & 0xffffffff, seed >> 32)
    ctr = (trial & 0xffffffff, trial >> 32, frame, (purpose << 16) | (draw << 8) | coord)
    purpose: 0 = feature noise, 1 = start-pose perturbation, 2 = random X0
One call yields r0..r3; ua = open52(r0 | r1 << 32), ub = open52(r2 | r3 << 32),
open52(x) = ((x >> 12) + 0.5) * 2**-52  in (0, 1)  (exact in FP64, never 0 or 1).
"""
import numpy as np

# noise.py:7-12
WHITE_NOISE, GAUSSIAN_MIXTURE, GAUSSIAN_BIMODAL, ALPHA_STABLE, UNIFORM = 1, 2, 3, 4, 5

PURPOSE_NOISE, PURPOSE_POSE, PURPOSE_X0 = 0, 1, 2

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  All inputs uint32
    arrays (broadcastable); returns four uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *[np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3, k0, k1)])
    c0, c1, c2, c3 = (a.astype(np.uint64) for a in (c0, c1, c2, c3))
    k0 = k0.astype(np.uint32).copy()
    k1 = k1.astype(np.uint32).copy()
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
        n1 = lo1
        n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        with np.errstate(over="ignore"):
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return tuple(a.astype(np.uint32) for a in (c0, c1, c2, c3))


def _open52(lo, hi):
    x = lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(32))
    return ((x >> np.uint64(12)).astype(np.float64) + 0.5) * (2.0 ** -52)


def uniforms(seed, trial, frame, coord, draw=0, purpose=PURPOSE_NOISE):
    """(ua, ub) in (0,1) for the given counter(s); arguments broadcast."""
    seed = int(seed)
    trial = np.asarray(trial, dtype=np.uint64)
    c0 = (trial & _MASK).astype(np.uint32)
    c1 = (trial >> np.uint64(32)).astype(np.uint32)
    c2 = np.asarray(frame, dtype=np.uint32)
    c3 = (np.uint32(purpose << 16) | np.uint32(draw << 8)) | np.asarray(coord, dtype=np.uint32)
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3,
                                   np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    return _open52(r0, r1), _open52(r2, r3)


def _box_muller(ua, ub):
    return np.sqrt(-2.0 * np.log(ua)) * np.cos(2.0 * np.pi * ub)


def sample(noise_type, params, seed, trial, frame, coord):
    """One fresh draw per (trial, frame, coord) — noise.py:120-207 per call.

    params: dict with std/mean/rho (Gaussian types) or alpha/beta/gamma/delta.
    Arguments broadcast; returns float64 array.
    """
    ua, ub = uniforms(seed, trial, frame, coord, 0)
    if noise_type == UNIFORM:
        return ua                                           # noise.py:120-123
    if noise_type == WHITE_NOISE:
        return params["std"] * _box_muller(ua, ub)          # noise.py:125-128
    if noise_type in (GAUSSIAN_MIXTURE, GAUSSIAN_BIMODAL):
        n = _box_muller(ua, ub)
        sel, _ = uniforms(seed, trial, frame, coord, 1)
        rho, mean, std = params["rho"], params["mean"], params["std"]
        if noise_type == GAUSSIAN_MIXTURE:                  # noise.py:130-138
            loc = np.where(sel > rho, 0.0, mean)
        else:                                               # noise.py:140-150
            loc = np.where(sel > rho, 0.0, np.where(sel > rho / 2, mean, -mean))
        return loc + std * n
    if noise_type == ALPHA_STABLE:                          # noise.py:152-207
        alpha, beta = params["alpha"], params["beta"]
        gamma, delta = params["gamma"], params["delta"]
        if alpha == 2:
            value = np.sqrt(2.0) * _box_muller(ua, ub)
        elif alpha == 1 and beta == 0:
            value = np.tan(-np.pi / 2 + np.pi * ua)
        elif alpha == 0.5 and abs(beta) == 1:
            value = beta / _box_muller(ua, ub) ** 2
        else:
            V = -np.pi / 2 + np.pi * ua
            W = -np.log(ub)
            if beta == 0:
                value = (np.sin(alpha * V) / np.cos(V) ** (1 / alpha)) * \
                    (np.cos(V * (1 - alpha)) / W) ** ((1 - alpha) / alpha)
            elif alpha != 1:
                constant = beta * np.tan(np.pi * alpha / 2)
                B = np.arctan(constant)
                S = (1 + constant ** 2) ** (1 / (2 * alpha))
                value = S * np.sin(alpha * V + B) / np.cos(V) ** (1 / alpha) * \
                    (np.cos((1 - alpha) * V - B) / W) ** ((1 - alpha) / alpha)
            else:
                sclshftV = np.pi / 2 + beta * V
                value = 2 / np.pi * (sclshftV * np.tan(V) - beta * np.log((W * np.cos(V)) / sclshftV))
        if alpha == 1:
            return gamma * value + (2 / np.pi) * beta * gamma * np.log(gamma) + delta
        return gamma * value + delta
    raise ValueError("noise type")


def noise_stream(noise_type, params, seed, trial, T, m, hold_cnt=0):
    """noise[T, m] for one trial with the pair-hold logic of noise.py:81-118.

    hold_cnt = noise_hold_cnt_max_predefined (0 when hold is off)."""
    frames = np.arange(T, dtype=np.uint32)[:, None]
    coords = np.arange(m, dtype=np.uint32)[None, :]
    fresh = np.broadcast_to(sample(noise_type, params, seed, np.uint64(trial), frames, coords), (T, m))
    return apply_hold(fresh, hold_cnt)


def apply_hold(fresh, hold_cnt=0):
    """noise.py:81-118 — fresh[k, j] is the value coordinate j draws IF its pair
    is redrawn at frame k; a pair holding an outlier (|x| > 20) is frozen for
    hold_cnt further frames."""
    T, m = fresh.shape
    out = np.zeros((T, m))
    cnt = np.zeros(m // 2)
    cnt_max = np.zeros(m // 2)
    cur = np.zeros(m)
    for k in range(T):
        for i in range(m // 2):
            if cnt[i] >= cnt_max[i]:
                cnt[i] = 0
                pair = fresh[k, 2 * i:2 * i + 2]
                outlier = bool(np.any(np.abs(pair) > 20))
                cnt_max[i] = hold_cnt if outlier else 0
                cur[2 * i:2 * i + 2] = pair
            else:
                cnt[i] += 1
        out[k] = cur
    return out


def start_pose(q_start, pose_seed, trial, change_q_start=True):
    """main.py:129-134 with the uniform draws taken from Philox."""
    q = np.array(q_start, dtype=float)
    if change_q_start:
        r, _ = uniforms(pose_seed, np.uint64(trial), 0, np.arange(2, dtype=np.uint32), 0, PURPOSE_POSE)
        q[0] = q_start[0] + 2 * (r[0] - 1) * (np.pi / 18)
        q[1] = q_start[1] + 2 * (r[1] - 1) * (np.pi / 9)
    return q


def random_x0(seed, trial, N):
    """experiment.py:117 (unseeded U(0,1)^N in the reference) on Philox."""
    u, _ = uniforms(seed, np.uint64(trial), 0, np.arange(N, dtype=np.uint32), 0, PURPOSE_X0)
    return u
