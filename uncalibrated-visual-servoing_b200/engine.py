"""Thin host layer over the C ABI: torch tensors in, torch tensors out.

torch is used for device memory, streams and (elsewhere) torch.distributed only;
every computation below is one call into libuvs_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

# disc centres of the three scenes, feature order red, green, blue, pink
TARGETS = {
    1: [[0.650, 0.000, 0.012]],
    3: [[0.575, -0.075, 0.012], [0.650, 0.000, 0.012], [0.575, 0.075, 0.012]],
    4: [[0.625, 0.100, 0.012], [0.700, 0.175, 0.012], [0.625, 0.250, 0.012], [0.550, 0.175, 0.012]],
}


def _require_cuda():
    if not torch.cuda.is_available():
        raise L.UvsError("uvs_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def dev(x, dtype=torch.float64, device=None):
    """Host array / tensor -> contiguous CUDA tensor of `dtype`."""
    _require_cuda()
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(np.ascontiguousarray(x))
    return x.to(device=device or "cuda", dtype=dtype).contiguous()


def _targets_arg(nf, targets):
    t = np.asarray(TARGETS[nf] if targets is None else targets, dtype=np.float64).reshape(-1)
    if t.size != 3 * nf:
        raise ValueError(f"targets must hold {nf} points")
    return (C.c_double * t.size)(*t)


def filter_params(method, nf, *, mode=L.MODE_AUTO, joseph=True, mckf_skip=True, kernel_bw=10.0,
                  fpi_threshold=0.1, fpi_epoch_max=1000, annealing=False, Q=None, R=None):
    """Builds the C struct; Q / R are CUDA tensors kept alive on the returned object."""
    p = L.FilterParams()
    p.method, p.nf, p.mode = int(method), int(nf), int(mode)
    p.joseph, p.mckf_skip = int(bool(joseph)), int(bool(mckf_skip))
    p.fpi_epoch_max, p.annealing = int(fpi_epoch_max), int(bool(annealing))
    p.kernel_bw, p.fpi_threshold = float(kernel_bw), float(fpi_threshold)
    p._keep = (dev(Q), dev(R))
    p.Q = _ptr(p._keep[0]).value
    p.R = _ptr(p._keep[1]).value
    return p


def noise_params(noise_type, *, seed=0, hold_cnt=0, std=0.0, mean=0.0, rho=0.0, alpha=2.0, beta=0.0, gamma=1.0,
                 delta=0.0):
    p = L.NoiseParams()
    p.type, p.hold_cnt = int(noise_type), int(hold_cnt)
    p.std, p.mean, p.rho = float(std), float(mean), float(rho)
    p.alpha, p.beta, p.gamma, p.delta = float(alpha), float(beta), float(gamma), float(delta)
    p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return p


def device_info():
    _require_cuda()
    vals = [C.c_int(0) for _ in range(4)]
    L.check(L.load().uvs_device_info(*[C.byref(v) for v in vals]), "uvs_device_info")
    return dict(sm_count=vals[0].value, cc=(vals[1].value, vals[2].value), sm_clock_khz=vals[3].value)


def filter_step(p, X, P, h, Z, sigma, *, p_layout=L.P_DENSE, want_K=False):
    """One predict+update of B filters in place (X [B,N], P [B,N,N] or [B,m,6,6])."""
    _require_cuda()
    B, m = Z.shape
    K = torch.empty((B, 6 * m, m), dtype=torch.float64, device=X.device) if want_K else None
    kappa = torch.empty((B, m), dtype=torch.float64, device=X.device)
    flags = torch.empty((B,), dtype=torch.int32, device=X.device)
    L.check(L.load().uvs_filter_step(C.byref(p), B, _ptr(X), _ptr(P), p_layout, _ptr(h), _ptr(Z), float(sigma),
                                     _ptr(K), _ptr(kappa), _ptr(flags), _stream()), "uvs_filter_step")
    return K, kappa, flags


def replay(p, f, dq, f_init, X, P=None, *, p_layout=L.P_DENSE, k_max=0, want_X=True, want_P=False, want_kappa=False):
    """T frames of B filters (f [B,T,m], dq [B,T,6]); X (and P) updated in place."""
    _require_cuda()
    B, T, m = f.shape
    N = 6 * m
    d = X.device
    X_out = torch.empty((B, T, N), dtype=torch.float64, device=d) if want_X else None
    if want_P:
        shape = (B, T, N, N) if p_layout == L.P_DENSE else (B, T, m, 6, 6)
        P_out = torch.zeros(shape, dtype=torch.float64, device=d)
    else:
        P_out = None
    kap = torch.empty((B, T, m), dtype=torch.float64, device=d) if want_kappa else None
    flags = torch.empty((B, T), dtype=torch.int32, device=d)
    L.check(L.load().uvs_replay(C.byref(p), B, T, int(k_max), _ptr(f), _ptr(dq), _ptr(f_init), _ptr(X), _ptr(P),
                                p_layout, _ptr(X_out), _ptr(P_out), _ptr(kap), _ptr(flags), _stream()), "uvs_replay")
    return X_out, P_out, kap, flags


def ibvs(nf, J, e, kappa, gain):
    _require_cuda()
    B = J.shape[0]
    dq = torch.empty((B, 6), dtype=torch.float64, device=J.device)
    flags = torch.empty((B,), dtype=torch.int32, device=J.device)
    L.check(L.load().uvs_ibvs(nf, B, _ptr(J), _ptr(e), _ptr(kappa), float(gain), _ptr(dq), _ptr(flags), _stream()),
            "uvs_ibvs")
    return dq, flags


def ur10_fk(q, want_jacobian=True):
    _require_cuda()
    B = q.shape[0]
    T06 = torch.empty((B, 4, 4), dtype=torch.float64, device=q.device)
    Jg = torch.empty((B, 6, 6), dtype=torch.float64, device=q.device) if want_jacobian else None
    L.check(L.load().uvs_ur10_fk(B, _ptr(q), _ptr(T06), _ptr(Jg), _stream()), "uvs_ur10_fk")
    return T06, Jg


def project(nf, q, targets=None):
    _require_cuda()
    B = q.shape[0]
    f = torch.empty((B, 2 * nf), dtype=torch.float64, device=q.device)
    depth = torch.empty((B, nf), dtype=torch.float64, device=q.device)
    L.check(L.load().uvs_project(nf, B, _ptr(q), _targets_arg(nf, targets), _ptr(f), _ptr(depth), _stream()),
            "uvs_project")
    return f, depth


def initial_guess(nf, q, targets=None):
    _require_cuda()
    B = q.shape[0]
    X0 = torch.empty((B, 12 * nf), dtype=torch.float64, device=q.device)
    f = torch.empty((B, 2 * nf), dtype=torch.float64, device=q.device)
    L.check(L.load().uvs_initial_guess(nf, B, _ptr(q), _targets_arg(nf, targets), _ptr(X0), _ptr(f), _stream()),
            "uvs_initial_guess")
    return X0, f


def noise_fill(np_, trial0, B, T, m, device="cuda"):
    _require_cuda()
    out = torch.empty((B, T, m), dtype=torch.float64, device=device)
    L.check(L.load().uvs_noise_fill(C.byref(np_), int(trial0), B, T, m, _ptr(out), _stream()), "uvs_noise_fill")
    return out


def start_poses(q_start, pose_seed, change_q_start, trial0, B, device="cuda"):
    _require_cuda()
    out = torch.empty((B, 6), dtype=torch.float64, device=device)
    qs = (C.c_double * 6)(*[float(v) for v in q_start])
    L.check(L.load().uvs_start_poses(qs, int(pose_seed) & 0xFFFFFFFFFFFFFFFF, int(bool(change_q_start)), int(trial0),
                                     B, _ptr(out), _stream()), "uvs_start_poses")
    return out


def closed_loop_cfg(filt, noise, *, T, dt, ibvs_gain, q_start, desired_f, targets=None, initial_guess=True,
                    change_q_start=True, fov_check=False, pose_seed=12345, sweep=None, sweep_epoch=1,
                    sweep_param=0, P0=None, v_max=0.0, a_max=0.0, servo_gain=0.0, trial_stride=1,
                    sensor_offset=None, disc_radius=0.0, servo_substeps=0, servo_kp=0.0):
    cfg = L.ClosedLoopCfg()
    cfg.filt = filt
    cfg.noise = noise
    cfg.T, cfg.initial_guess = int(T), int(bool(initial_guess))
    cfg.change_q_start, cfg.fov_check = int(bool(change_q_start)), int(bool(fov_check))
    sweep = [] if sweep is None else [float(v) for v in sweep]
    if len(sweep) > L.UVS_MAX_SWEEP:
        raise ValueError("sweep too long")
    cfg.sweep_len, cfg.sweep_epoch, cfg.sweep_param = len(sweep), int(sweep_epoch), int(sweep_param)
    cfg.trial_stride = int(trial_stride)        # batch entry b = global trial trial0 + b * trial_stride (interleaved shards)
    for i, v in enumerate(sweep):
        cfg.sweep[i] = v
    cfg.dt, cfg.ibvs_gain = float(dt), float(ibvs_gain)
    for i in range(6):
        cfg.q_start[i] = float(q_start[i])
    m = 2 * filt.nf
    for i in range(m):
        cfg.desired_f[i] = float(desired_f[i])
    t = np.asarray(TARGETS[filt.nf] if targets is None else targets, dtype=np.float64).reshape(-1)
    for i in range(3 * filt.nf):
        cfg.targets[i] = float(t[i])
    cfg.pose_seed = int(pose_seed) & 0xFFFFFFFFFFFFFFFF
    cfg._keep = (filt, dev(P0))
    cfg.P0 = _ptr(cfg._keep[1]).value
    cfg.v_max, cfg.a_max, cfg.servo_gain = float(v_max), float(a_max), float(servo_gain)   # servo model (0 = kinematic plant)
    # scene plant (include/uvs_b200.h): sensor mount offset, finite discs, sub-stepped position servo; all 0 = off
    for i, v in enumerate(sensor_offset if sensor_offset is not None else (0.0, 0.0, 0.0)):
        cfg.sensor_offset[i] = float(v)
    cfg.disc_radius, cfg.servo_kp, cfg.servo_substeps = float(disc_radius), float(servo_kp), int(servo_substeps)
    return cfg


def _loop_outputs(cfg, B, want_log, want_state, want_trace, p_layout, device):
    m = 2 * cfg.filt.nf
    N = 6 * m
    out = dict(stats=torch.empty((B, 3), dtype=torch.float64, device=device),
               status=torch.empty((B,), dtype=torch.int32, device=device),
               frames=torch.empty((B,), dtype=torch.int32, device=device),
               flags=torch.empty((B,), dtype=torch.int32, device=device))
    if want_log:
        out["log"] = torch.zeros((B, cfg.T, 6 + 2 * m), dtype=torch.float64, device=device)
    pshape = (N, N) if p_layout == L.P_DENSE else (m, 6, 6)
    if want_state:
        out["X"] = torch.empty((B, N), dtype=torch.float64, device=device)
        out["P"] = torch.zeros((B,) + pshape, dtype=torch.float64, device=device)
    if want_trace:          # per-frame trace (uvs_closed_loop_io, ABI 4); frames a trial did not run stay zero
        out["X_log"] = torch.zeros((B, cfg.T, N), dtype=torch.float64, device=device)
        out["P_log"] = torch.zeros((B, cfg.T, m, 6, 6), dtype=torch.float64, device=device)     # diagonal blocks
        out["dq_log"] = torch.zeros((B, cfg.T, 6), dtype=torch.float64, device=device)
        out["kappa_log"] = torch.zeros((B, cfg.T, m), dtype=torch.float64, device=device)
        out["flags_log"] = torch.zeros((B, cfg.T), dtype=torch.int32, device=device)
    return out


def _loop_io(out, p_layout, noise_in, q0_in, X0_in, dq_in):
    io = L.ClosedLoopIO()
    io.stats, io.status = _ptr(out["stats"]).value, _ptr(out["status"]).value
    io.frames, io.flags = _ptr(out.get("frames")).value, _ptr(out.get("flags")).value
    io.log, io.X_out, io.P_out = _ptr(out.get("log")).value, _ptr(out.get("X")).value, _ptr(out.get("P")).value
    io.noise_in, io.q0_in, io.X0_in = _ptr(noise_in).value, _ptr(q0_in).value, _ptr(X0_in).value
    io.p_layout = int(p_layout)
    io.dq_in = _ptr(dq_in).value
    io.X_log, io.P_log, io.dq_log = _ptr(out.get("X_log")).value, _ptr(out.get("P_log")).value, _ptr(out.get("dq_log")).value
    io.kappa_log, io.flags_log = _ptr(out.get("kappa_log")).value, _ptr(out.get("flags_log")).value
    return io


def closed_loop(cfg, trial0, B, *, want_log=False, want_state=False, want_trace=False, p_layout=L.P_BLOCKS,
                noise_in=None, q0_in=None, X0_in=None, dq_in=None, device="cuda", out=None):
    """Runs B closed-loop trials; returns a dict of CUDA tensors (no host sync).  ``want_trace`` adds the per-frame
    X / P / dq / kappa / flags logs; ``dq_in`` [B,T,6] replays a recorded command stream (the replay gate of the
    closed-loop kernels, include/uvs_b200.h)."""
    _require_cuda()
    if out is None:
        out = _loop_outputs(cfg, B, want_log, want_state, want_trace, p_layout, device)
    io = _loop_io(out, p_layout, noise_in, q0_in, X0_in, dq_in)
    L.check(L.load().uvs_closed_loop(C.byref(cfg), int(trial0), int(B), C.byref(io), _stream()), "uvs_closed_loop")
    return out


def twist_cfg(*, x0_mode=L.X0_IMAGE_JACOBIAN, depth_mode=L.DEPTH_DISTANCE, dp_zero_mask=0, focal_x0=1024.0,
              focal=200.91975387103875, err_threshold=-1.0, stale_fk=False):
    """The parameters only the reference's tests/ scripts have (include/uvs_b200.h: uvs_twist_cfg)."""
    tw = L.TwistCfg()
    tw.x0_mode, tw.depth_mode, tw.dp_zero_mask = int(x0_mode), int(depth_mode), int(dp_zero_mask)
    tw.stale_fk = int(bool(stale_fk))
    tw.focal_x0, tw.focal, tw.err_threshold = float(focal_x0), float(focal), float(err_threshold)
    return tw


def twist_loop(cfg, tw, trial0, B, *, want_log=False, want_state=False, want_trace=False, p_layout=L.P_DENSE,
               noise_in=None, q0_in=None, X0_in=None, device="cuda", out=None):
    """B trials of the tests/-style camera-twist loop (tests/{ibvs,kalman,mckf}_*.py); same outputs as
    ``closed_loop``.  cfg.ibvs_gain is the scripts' GAIN, cfg.dt their TS."""
    _require_cuda()
    if out is None:
        out = _loop_outputs(cfg, B, want_log, want_state, want_trace, p_layout, device)
    io = _loop_io(out, p_layout, noise_in, q0_in, X0_in, None)
    L.check(L.load().uvs_twist_loop(C.byref(cfg), C.byref(tw), int(trial0), int(B), C.byref(io), _stream()),
            "uvs_twist_loop")
    return out


_pinned = {}


def _host_result_buffers(B):
    """Page-locked result buffers, cached per batch size (pinning is the expensive part), so the device->host
    copies of ``uvs_closed_loop_host`` are plain DMAs."""
    if B not in _pinned:
        _pinned.clear()
        _pinned[B] = (torch.empty((B, 3), dtype=torch.float64, pin_memory=True),
                      torch.empty((B,), dtype=torch.int32, pin_memory=True),
                      torch.empty((B,), dtype=torch.int32, pin_memory=True))
    return _pinned[B]


def closed_loop_host(cfg, trial0, B, pinned=True):
    """The host-buffer entry point: numpy arrays out, synchronous.  With ``pinned`` the arrays are views of cached
    page-locked buffers that the next call of the same size overwrites (copy them to keep them)."""
    _require_cuda()
    if pinned and B > 0:
        stats, status, frames = (t.numpy() for t in _host_result_buffers(B))
    else:
        stats = np.empty((B, 3), dtype=np.float64)
        status = np.empty((B,), dtype=np.int32)
        frames = np.empty((B,), dtype=np.int32)
    L.check(L.load().uvs_closed_loop_host(C.byref(cfg), int(trial0), int(B),
                                          stats.ctypes.data_as(C.POINTER(C.c_double)),
                                          status.ctypes.data_as(C.POINTER(C.c_int32)),
                                          frames.ctypes.data_as(C.POINTER(C.c_int32))), "uvs_closed_loop_host")
    return stats, status, frames


def closed_loop_stream(cfg, trial0, B, chunk=4096, sink=None):
    """Streams the trajectory log of B closed-loop trials to the host in chunks (uvs_closed_loop_stream: two device and
    two page-locked host buffers, the copy of chunk i overlaps the launch of chunk i + 1).  ``sink(first_trial, log, stats,
    status, frames)`` receives numpy VIEWS of the page-locked buffers (log [n,T,6+2m]), valid during the call only; a
    truthy return value stops the stream.  Returns dict(d2h_gbs, wall_s)."""
    _require_cuda()
    m = 2 * cfg.filt.nf
    T, W = cfg.T, 6 + 2 * m
    err = []

    def _sink(user, first, n, log, stats, status, frames):
        try:
            n = int(n)
            stop = sink(int(first), np.ctypeslib.as_array(log, shape=(n, T, W)), np.ctypeslib.as_array(stats, shape=(n, 3)),
                        np.ctypeslib.as_array(status, shape=(n,)), np.ctypeslib.as_array(frames, shape=(n,)))
            return 1 if stop else 0
        except BaseException as e:           # never let an exception cross the C boundary
            err.append(e)
            return 1

    cb = L.LOG_SINK(_sink) if sink is not None else None
    gbs, wall = C.c_double(0), C.c_double(0)
    rc = L.load().uvs_closed_loop_stream(C.byref(cfg), int(trial0), int(B), int(chunk),
                                         C.cast(cb, C.c_void_p) if cb is not None else None, None,
                                         C.byref(gbs), C.byref(wall))
    if err:
        raise err[0]
    L.check(rc, "uvs_closed_loop_stream")
    return dict(d2h_gbs=gbs.value, wall_s=wall.value)


def reduce_stats(stats, status, trial0, sweep_epoch, n_levels, out=None, trial_stride=1):
    """Per-level sums [n_levels,3,3] = {n_ok, sum, sum of squares} and n_fail [n_levels].  ``out`` = (sums, n_fail)
    to ACCUMULATE into existing buffers (e.g. two views of one packed tensor that is all-reduced once)."""
    _require_cuda()
    B = stats.shape[0]
    if out is not None:
        sums, n_fail = out
    else:
        sums = torch.zeros((n_levels, 3, 3), dtype=torch.float64, device=stats.device)
        n_fail = torch.zeros((n_levels,), dtype=torch.float64, device=stats.device)
    L.check(L.load().uvs_reduce_stats(B, int(trial0), int(trial_stride), int(sweep_epoch), int(n_levels), _ptr(stats), _ptr(status),
                                      _ptr(sums), _ptr(n_fail), _stream()), "uvs_reduce_stats")
    return sums, n_fail


def fp64_peak(iters=4096):
    _require_cuda()
    tf, ms = C.c_double(0), C.c_double(0)
    L.check(L.load().uvs_fp64_peak(int(iters), C.byref(tf), C.byref(ms)), "uvs_fp64_peak")
    return tf.value, ms.value
