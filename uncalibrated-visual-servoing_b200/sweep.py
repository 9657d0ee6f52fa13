"""Batch driver: the reference's ``main.py`` sweep (main.py:104-207) as ONE sharded GPU job, plus the
statistics of ``results/plot_errorbar.m:39-98`` and the ``results.csv`` writer (main.py:152-196).

The reference loops ``for rho in rho_list: for i in range(epoch): Experiment(...).run()`` serially with
CoppeliaSim in the loop.  Here experiment ``k`` of that queue is global trial ``k`` of one closed-loop
launch: its noise level is ``rho_list[k // epoch]`` (main.py:121-127), its start pose the main.py:129-134
perturbation drawn from the Philox stream keyed by ``experiments.seed``, its noise the Philox stream keyed
by ``noise.seed`` and the trial id (the reference's ``seed = seed + 1`` per experiment, main.py:138-139).

Multi-GPU: trials are independent, so rank r of W runs the contiguous range
``shard_range(n, r, W)`` and nothing is exchanged during the run.  One all-reduce of the per-level
``{n_ok, sum x, sum x^2}`` block (+ ``n_fail``) gives mean / std everywhere; medians need the per-trial
values, which ``gather_trials`` collects on request.  The reduction helpers work on whatever device the
tensors live on, so the N>1 path is tested on CPU with the gloo backend (tests/test_sweep_host.py); the
closed-loop launch itself needs the CUDA library.
"""
import csv
import json
import logging
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

STAT_NAMES = ("ise", "iae", "itae")

# column order of main.py:152-194 for m = 8; other feature counts keep the same layout with m columns
def csv_columns(m=8):
    return (["experiment_id", "status", "rho", "t"] + [f"q_{i + 1}" for i in range(6)] +
            ["camera_x", "camera_y", "camera_z", "camera_roll", "camera_pitch", "camera_yaw"] +
            [f"f_{i + 1}" for i in range(m)] + [f"desired_f_{i + 1}" for i in range(m)] +
            [f"noise_{i + 1}" for i in range(m)] + ["kernel_bw"])


# ------------------------------------------------------------------------------------- config
class ConfigError(SystemExit):
    """main.py:99-101,71-73,88-90: a bad config is fatal (``sys.exit()``)."""


def load_config(config):
    """``config`` = path to a config.json or the parsed dict.  Returns the flat run description.
    Key names and failure behaviour follow main.py:16-101 (missing key / unknown enum name -> exit)."""
    from .experiment import Method
    from .noise import NoiseType
    if isinstance(config, (str, os.PathLike)):
        with open(config, "r", encoding="utf-8") as fh:
            config = json.load(fh)
    log = logging.getLogger(__name__)
    try:
        ex, est, nz = config["experiments"], config["estimator"], config["noise"]
        run = dict(dt=ex["dt"], t_max=ex["t_max"], epoch=int(ex["epoch"]), ibvs_gain=ex["ibvs_gain"],
                   q_start=[float(v) for v in ex["q_start"]], desired_f=[float(v) for v in ex["desired_f"]],
                   visualization=ex["visualization"], change_q_start=bool(ex["change_q_start"]),
                   experiment_seed=ex["seed"])
        if est["method"] not in Method.__members__:
            log.critical("Estimation method " + str(est["method"]) + " unknown.")
            raise ConfigError(1)
        run["method"] = Method[est["method"]]
        run["method_params"] = dict(est["estimator_params"])
        # additive, optional: "plant": {"servo_gain", "v_max", "a_max", "sensor_offset", "disc_radius", "servo_substeps",
        # "servo_kp", "fov_check"} (include/uvs_b200.h: the servo limits and the scene plant); absent = the
        # kinematic plant q <- q + dq dt of experiment.py:320,335
        for key, val in config.get("plant", {}).items():
            if key not in ("servo_gain", "v_max", "a_max", "sensor_offset", "disc_radius", "servo_substeps", "servo_kp",
                           "fov_check"):
                log.critical("Plant parameter " + str(key) + " unknown.")
                raise ConfigError(1)
            run["method_params"][key] = val if isinstance(val, (list, tuple, bool)) else float(val)
        if nz["type"] not in NoiseType.__members__:
            log.critical("Noise type " + str(nz["type"]) + " unknown.")
            raise ConfigError(1)
        run["noise_type"] = NoiseType[nz["type"]]
        run["noise_params"] = dict(nz["noise_params"])
        run["noise_hold"], run["noise_hold_time"], run["noise_seed"] = nz["hold"], nz["hold_time"], nz["seed"]
        run["log_level"] = config.get("log_level", "INFO")
    except KeyError as e:
        log.critical("Could not load config key: " + str(e))
        raise ConfigError(1)
    # main.py:104-106
    run["sweep_param"] = "alpha" if run["noise_type"].name == "ALPHA_STABLE" else "rho"
    run["rho_list"] = [float(v) for v in (np.linspace(1, 2, 12) if run["sweep_param"] == "alpha"
                                          else np.linspace(0, 0.2, 12))]
    return run


# ------------------------------------------------------------------------------------- sharding
def shard_range(n_trials, rank, world):
    """Contiguous trial range [lo, hi) of ``rank``; the first ``n % world`` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    base, extra = divmod(int(n_trials), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_interleaved(n_trials, rank, world):
    """(first trial, stride, count) of ``rank`` when the trials are dealt round-robin: rank r runs r, r + W, r + 2W, ...
    Every rank then sees the same mix of noise levels — a frame's cost depends on the level (alpha = 1 and alpha = 2 use
    cheap samplers), so contiguous shards of a sweep do not finish together (bench.py at 2 GPUs: 2.7 % apart)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside the group")
    n = int(n_trials)
    return rank, world, (n - rank + world - 1) // world if n > rank else 0


def _dist_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


# ------------------------------------------------------------------------------------- statistics
def level_sums_host(stats, status, trial0, epoch, n_levels):
    """numpy/torch (any device) version of ``uvs_reduce_stats`` for the host-side tests and for CPU
    tensors: sums [n_levels,3,3] = {n_ok, sum, sum sq} per level and statistic, n_fail [n_levels]."""
    stats = torch.as_tensor(stats, dtype=torch.float64)
    status = torch.as_tensor(status)
    B = stats.shape[0]
    lvl = torch.clamp((torch.arange(B, device=stats.device) + int(trial0)) // max(int(epoch), 1), max=n_levels - 1)
    ok = (status == 0)
    sums = torch.zeros((n_levels, 3, 3), dtype=torch.float64, device=stats.device)
    n_fail = torch.zeros((n_levels,), dtype=torch.float64, device=stats.device)
    okf = ok.to(torch.float64)
    x = torch.where(ok[:, None], stats, torch.zeros_like(stats))
    sums[:, :, 0].index_add_(0, lvl, okf[:, None].expand(B, 3).contiguous())
    sums[:, :, 1].index_add_(0, lvl, x)
    sums[:, :, 2].index_add_(0, lvl, x * x)
    n_fail.index_add_(0, lvl, 1.0 - okf)
    return sums, n_fail


def merge_level_sums(sums, n_fail, group=None):
    """The one collective of the job: all-reduce(sum) of the per-level block (NCCL over NVLink on GPUs,
    gloo in the CPU tests).  In place; returns the tensors."""
    _, world = _dist_info(group)
    if world > 1:
        packed = torch.cat([sums.reshape(-1), n_fail.reshape(-1)])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        sums.copy_(packed[:sums.numel()].view_as(sums))
        n_fail.copy_(packed[sums.numel():].view_as(n_fail))
    return sums, n_fail


def gather_trials(stats, status, group=None):
    """All ranks' per-trial rows in trial order (shards are contiguous and rank-ordered)."""
    rank, world = _dist_info(group)
    if world == 1:
        return stats, status
    counts = [torch.zeros(1, dtype=torch.int64, device=stats.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([stats.shape[0]], dtype=torch.int64, device=stats.device), group=group)
    nmax = int(max(c.item() for c in counts))
    pad = torch.zeros((nmax, 4), dtype=torch.float64, device=stats.device)
    pad[:stats.shape[0], :3] = stats
    pad[:stats.shape[0], 3] = status.to(torch.float64)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    rows = torch.cat([p[:int(c.item())] for p, c in zip(parts, counts)])
    return rows[:, :3].contiguous(), rows[:, 3].to(torch.int32)


def level_summary(sums, n_fail, levels=None):
    """mean / std (N-1, MATLAB ``std``) per level from the reduced sums — plot_errorbar.m:87-94."""
    s = torch.as_tensor(sums).detach().cpu().numpy()
    nf = torch.as_tensor(n_fail).detach().cpu().numpy()
    out = []
    for i in range(s.shape[0]):
        n = s[i, 0, 0]
        row = {"level": None if levels is None else float(levels[i]), "n_ok": int(round(n)), "n_fail": int(round(nf[i]))}
        for j, name in enumerate(STAT_NAMES):
            sx, sxx = s[i, j, 1], s[i, j, 2]
            row[name + "_mean"] = float(sx / n) if n > 0 else float("nan")
            var = (sxx - sx * sx / n) / (n - 1) if n > 1 else float("nan")
            row[name + "_std"] = float(np.sqrt(max(var, 0.0))) if n > 1 else float("nan")
        out.append(row)
    return out


def level_quantiles(stats, status, epoch, n_levels, q=(0.25, 0.5, 0.75)):
    """Per-level quantiles over SUCCESS trials (median of plot_errorbar.m:90, box of plot_boxchart.m)."""
    stats = np.asarray(torch.as_tensor(stats).detach().cpu())
    status = np.asarray(torch.as_tensor(status).detach().cpu())
    lvl = np.minimum(np.arange(len(stats)) // max(int(epoch), 1), n_levels - 1)
    out = np.full((n_levels, 3, len(q)), np.nan)
    for i in range(n_levels):
        sel = stats[(lvl == i) & (status == 0)]
        if len(sel):
            out[i] = np.quantile(sel, q, axis=0).T
    return out


# ------------------------------------------------------------------------------------- CSV
def results_frame(experiment_ids, status, rho, logs, frames, desired_f, dt, kernel_bw_log=None):
    """The rows main.py:152-194 appends for the given trials, as ONE pandas DataFrame built column-wise (no Python
    loop over rows): experiment constants repeated on every frame row, ``status`` spelled like the enum's ``str()``.

    logs: [n, T, 6 + 2m] (q, f, noise) as produced by the closed-loop kernels; frames[n] = rows to emit per trial."""
    import pandas as pd
    from .experiment import ExperimentStatus
    from .ur10_simulation import camera_log_from_q
    logs = np.asarray(logs, dtype=float)
    n, T, W = logs.shape
    m = (W - 6) // 2
    frames = np.asarray(frames, dtype=np.int64)
    keep = (np.arange(T)[None, :] < frames[:, None])                       # [n,T] rows a trial actually ran
    rows = logs[keep]                                                      # [R, W], trial-major then frame order
    trial_of = np.repeat(np.arange(n), frames)
    frame_of = np.nonzero(keep)[1]
    names = {int(e.value): str(e) for e in ExperimentStatus}
    data = {"experiment_id": np.asarray(experiment_ids, dtype=np.int64)[trial_of],
            "status": np.array([names[int(v)] for v in np.asarray(status)], dtype=object)[trial_of],
            "rho": np.asarray(rho, dtype=float)[trial_of],
            "t": (frame_of + 1) * float(dt)}
    cam = camera_log_from_q(rows[:, 0:6])                                  # experiment.py:324, one batched FK launch
    for i in range(6):
        data[f"q_{i + 1}"] = rows[:, i]
    for i, name in enumerate(("camera_x", "camera_y", "camera_z", "camera_roll", "camera_pitch", "camera_yaw")):
        data[name] = cam[:, i]
    for i in range(m):
        data[f"f_{i + 1}"] = rows[:, 6 + i]
    for i in range(m):
        data[f"desired_f_{i + 1}"] = np.full(len(rows), float(desired_f[i]))
    for i in range(m):
        data[f"noise_{i + 1}"] = rows[:, 6 + m + i]
    if kernel_bw_log is None:
        data["kernel_bw"] = -np.ones(len(rows))
    else:
        data["kernel_bw"] = np.asarray(kernel_bw_log, dtype=float)[:, :T][keep]
    return pd.DataFrame(data, columns=csv_columns(m))


def write_results_csv(path, experiment_ids, status, rho, logs, frames, desired_f, dt, kernel_bw_log=None,
                      header=True, append=False):
    """Appends the frame logs of the given trials in the reference schema (main.py:152-196), through the same
    ``DataFrame.to_csv(mode='a', index=False)`` call the reference makes — so numbers are formatted identically."""
    df = results_frame(experiment_ids, status, rho, logs, frames, desired_f, dt, kernel_bw_log)
    df.to_csv(path, mode="a" if append else "w", index=False, header=header)
    return len(df)


def stream_results(cfg, lo, hi, levels, epoch, csv_path=None, npy_path=None, chunk=4096, kernel_bw_row=None,
                   desired_f=None, dt=0.05, header=True):
    """Runs trials [lo, hi) once with the trajectory log STREAMED to the host (engine.closed_loop_stream) and consumed
    chunk by chunk while the GPU computes the next chunk: rows appended to ``csv_path`` in the reference schema and/or
    the raw [n,T,6+2m] log written to ``npy_path`` (a .npy memory map — the format for jobs whose CSV would be
    hundreds of GB).  Returns (stats [n,3], status [n], frames [n], info) with the streaming rates."""
    from . import engine
    n = hi - lo
    m = 2 * cfg.filt.nf
    stats = np.zeros((n, 3)); status = np.zeros(n, dtype=np.int32); frames = np.zeros(n, dtype=np.int32)
    mm = None
    if npy_path is not None:
        mm = np.lib.format.open_memmap(npy_path, mode="w+", dtype=np.float64, shape=(n, cfg.T, 6 + 2 * m))
    state = {"header": header, "rows": 0, "csv_s": 0.0}

    def sink(first, log, st, ss, fr):
        import time
        a, b = first - lo, first - lo + len(st)
        stats[a:b], status[a:b], frames[a:b] = st, ss, fr
        if mm is not None:
            mm[a:b] = log
        if csv_path is not None:
            t0 = time.perf_counter()
            ids = np.arange(first, first + len(st))
            lvl = np.minimum(ids // max(int(epoch), 1), len(levels) - 1)
            bw = None if kernel_bw_row is None else np.tile(kernel_bw_row, (len(st), 1))
            state["rows"] += write_results_csv(csv_path, ids, ss, np.asarray(levels)[lvl], log, fr, desired_f, dt,
                                               kernel_bw_log=bw, header=state["header"], append=not state["header"])
            state["header"] = False
            state["csv_s"] += time.perf_counter() - t0
        return 0

    info = engine.closed_loop_stream(cfg, lo, n, chunk=chunk, sink=sink)
    if mm is not None:
        mm.flush()
    info.update(rows=state["rows"], csv_s=state["csv_s"], log_bytes=int(frames.sum()) * (6 + 2 * m) * 8)
    return stats, status, frames, info


# ------------------------------------------------------------------------------------- the sweep
def run_sweep(config, *, epoch=None, csv_path=None, csv_trials=0, want_trials=False, mode=None, group=None,
              logger=None, stream_chunk=4096):
    """Runs the whole main.py queue (12 levels x epoch experiments) on this process group's GPUs.

    Returns, on every rank: ``levels`` (rho / alpha values), ``summary`` (per-level mean/std, counts),
    and with ``want_trials`` the gathered per-trial ``stats``/``status`` + ``quantiles``.  With
    ``csv_path`` rank 0 writes the frame logs of the first ``csv_trials`` experiments of every level
    (all of them when ``csv_trials`` is None) in the reference's results.csv schema."""
    from . import _lib as L
    from . import engine
    from .experiment import BatchExperiment, Method
    from .noise import NoiseProfiler
    run = load_config(config)
    log = logger or logging.getLogger(__name__)
    if epoch is not None:
        run["epoch"] = int(epoch)
    levels, ep = run["rho_list"], run["epoch"]
    n_trials = len(levels) * ep
    rank, world = _dist_info(group)
    lo, hi = shard_range(n_trials, rank, world)
    nparams = dict(run["noise_params"])
    nparams.setdefault("alpha" if run["sweep_param"] == "alpha" else "rho", levels[0])
    prof = NoiseProfiler(num_features=len(run["desired_f"]), noise_type=run["noise_type"], seed=run["noise_seed"],
                         noise_hold=run["noise_hold"], noise_hold_cnt=int(run["noise_hold_time"] / run["dt"]),
                         noise_params=nparams)
    mp = dict(run["method_params"])
    if mode is not None:
        mp["mode"] = mode
    exp = BatchExperiment(run["q_start"], run["desired_f"], prof, run["dt"], run["t_max"], run["ibvs_gain"],
                          run["method"], n_trials=hi - lo, trial0=lo, change_q_start=run["change_q_start"],
                          pose_seed=run["experiment_seed"] if run["experiment_seed"] is not None else 0,
                          sweep=levels, sweep_epoch=ep, sweep_param=run["sweep_param"], method_params=mp)
    log.info("Experiments %d..%d of %d on rank %d/%d", lo, hi, n_trials, rank, world)
    T = int(run["t_max"] / run["dt"])
    bw_row = None
    if run["method"] == Method.MCKF:                        # experiment.py:330: only MCKF logs its bandwidth
        kk = np.arange(T)
        bw_row = mp["kernel_bw"] + (100.0 * (1 - kk / T) if mp.get("annealing") else 0.0 * kk)
    streamed = csv_path is not None and csv_trials is None
    stream_info = None
    if streamed:
        # main.py:152-196 logs EVERY experiment: one pass, the log streamed through page-locked buffers while the
        # next chunk computes; every rank writes its own contiguous part, rank 0 stitches them in rank order
        part = csv_path if world == 1 else f"{csv_path}.part{rank}"
        st_h, ss_h, fr_h, stream_info = stream_results(exp.config(), lo, hi, levels, ep, csv_path=part, chunk=stream_chunk,
                                                       kernel_bw_row=bw_row, desired_f=run["desired_f"], dt=run["dt"],
                                                       header=(rank == 0))
        out = {"stats": engine.dev(st_h), "status": engine.dev(ss_h, dtype=torch.int32)}
        if world > 1:
            dist.barrier(group)
            if rank == 0:
                with open(csv_path, "wb") as dst:
                    for r in range(world):
                        with open(f"{csv_path}.part{r}", "rb") as src:
                            while True:
                                buf = src.read(1 << 24)
                                if not buf:
                                    break
                                dst.write(buf)
                        os.remove(f"{csv_path}.part{r}")
    else:
        out = exp.run(to_host=False)
    sums, n_fail = engine.reduce_stats(out["stats"], out["status"], lo, ep, len(levels))
    merge_level_sums(sums, n_fail, group)
    result = {"levels": levels, "sweep_param": run["sweep_param"], "epoch": ep, "n_trials": n_trials,
              "summary": level_summary(sums, n_fail, levels), "shard": (lo, hi)}
    if want_trials:
        st, ss = gather_trials(out["stats"], out["status"], group)
        result["stats"], result["status"] = st.cpu().numpy(), ss.cpu().numpy()
        result["quantiles"] = level_quantiles(result["stats"], result["status"], ep, len(levels))
    if stream_info is not None:
        result["stream"] = stream_info
    if csv_path is not None and rank == 0 and not streamed:
        per = min(int(csv_trials), ep)
        ids = [lv * ep + i for lv in range(len(levels)) for i in range(per)]
        header = True
        for lv in range(len(levels)):                       # one small logged launch per level: bounded memory
            chunk = [k for k in ids if k // ep == lv]
            if not chunk:
                continue
            e2 = BatchExperiment(run["q_start"], run["desired_f"], prof, run["dt"], run["t_max"], run["ibvs_gain"],
                                 run["method"], n_trials=len(chunk), trial0=chunk[0],
                                 change_q_start=run["change_q_start"], pose_seed=exp.pose_seed, sweep=levels,
                                 sweep_epoch=ep, sweep_param=run["sweep_param"], method_params=mp)
            o = e2.run(want_log=True)
            bw = None if bw_row is None else np.tile(bw_row, (len(chunk), 1))
            write_results_csv(csv_path, chunk, o["status"], [levels[lv]] * len(chunk), o["log"], o["frames"],
                              run["desired_f"], run["dt"], kernel_bw_log=bw, header=header, append=not header)
            header = False
    ok = sum(r["n_ok"] for r in result["summary"])
    fail = sum(r["n_fail"] for r in result["summary"])
    log.info("Experiments status summary: %d fail | %d success", fail, ok)       # main.py:207
    return result


def main(argv=None):
    """``python main.py``-equivalent: reads config.json, runs the sweep, writes results.csv + summary.json
    under ./results/data/<timestamp>/ (main.py:12-13,101-103,196)."""
    import argparse
    import time
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--config", default=None,
                    help="config.json (default: ./config.json as main.py:16 reads it, else the copy shipped beside this "
                         "package, which equals the reference's)")
    ap.add_argument("--out", default=None, help="output directory (default results/data/<timestamp>)")
    ap.add_argument("--epoch", type=int, default=None, help="override experiments.epoch")
    ap.add_argument("--csv-trials", type=int, default=None,
                    help="experiments per level whose frame logs go to results.csv (default: all, as main.py)")
    ap.add_argument("--no-csv", action="store_true")
    a = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rank, _ = _dist_info()
    out_dir = a.out or os.path.join("results", "data", time.strftime("%d_%m_%Y_%H_%M_%S", time.localtime()))
    if rank == 0:
        os.makedirs(out_dir, exist_ok=True)
    logger = logging.getLogger("uvs_b200.main")
    logger.setLevel(logging.DEBUG)
    if not logger.handlers:
        ch = logging.StreamHandler(sys.stdout)
        ch.setFormatter(logging.Formatter('[%(levelname)s] %(asctime)s - %(name)s - %(message)s'))
        logger.addHandler(ch)
    cfg_path = a.config
    if cfg_path is None:
        here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config.json")
        cfg_path = "config.json" if os.path.exists("config.json") else here
    with open(cfg_path, "r", encoding="utf-8") as fh:
        cfg = json.load(fh)
    logger.handlers[0].setLevel(getattr(logging, str(cfg.get("log_level", "INFO")), logging.INFO))
    if rank == 0:
        with open(os.path.join(out_dir, "config.json"), "w", encoding="utf-8") as fh:
            json.dump(cfg, fh)
    res = run_sweep(cfg, epoch=a.epoch, csv_path=None if a.no_csv else os.path.join(out_dir, "results.csv"),
                    csv_trials=a.csv_trials, want_trials=True, logger=logger)
    if rank == 0:
        with open(os.path.join(out_dir, "summary.json"), "w", encoding="utf-8") as fh:
            json.dump({"levels": res["levels"], "sweep_param": res["sweep_param"], "epoch": res["epoch"],
                       "summary": res["summary"], "quantiles": res["quantiles"].tolist()}, fh, indent=1)
        logger.info("wrote %s", out_dir)
    if world > 1:
        dist.destroy_process_group()
    return res
