"""bench.py — RMCKF 4-feature closed-loop filter-steps/s (BASELINE.json metric) on N B200s.

    python bench.py --gpus 1 --steps 3 --warmup 3                  # this framework
    python bench.py --impl reference --steps 1 --warmup 0          # reference algorithm on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU

Workload: RMCKF (= Method.GMCKF), 4 features (48 states), closed-loop kinematic UR10 IBVS, impulsive
alpha-stable noise swept over alpha = linspace(1, 2, 12) with annealing, T = 300 frames.
  N = 1: BASELINE.json configs[2], 100 000 trials per step.
  N > 1: BASELINE.json configs[3], 1 000 000 trials per step sharded over the ranks (contiguous trial
         ranges, sweep.shard_range; strong scaling), unless --trials asks for a fixed count per GPU.
One "step" = one pass of the hot path over one batch = one kernel launch that runs every trial of the
shard through all T frames + the per-level statistics reduction into a device-resident block.  There is
no collective and no host synchronisation inside the step loop: the per-level sums of all steps are
all-reduced ONCE, packed, after the last step (inside the timed region).  A filter-step is one frame of
one trial; only frames actually executed are counted.

Prints ONE JSON line (rank 0).  `value` is timed with CUDA events with everything resident on
the device; `e2e` goes through the host-buffer C-ABI entry point (uvs_closed_loop_host:
launch, D2H of the results, synchronisation) per step.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

T_FRAMES = 300
DT = 0.05
Q_START = [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0]
DESIRED_F = [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0]
SWEEP = [float(v) for v in np.linspace(1.0, 2.0, 12)]
NOISE_SEED, POSE_SEED = 123456, 12345
# algorithmic FP64 flops of one RMCKF update with dense X, P, K (m = 8, N = 48)
M_, N_ = 8, 48
F_ALG = 4 * N_ ** 2 + 4 * N_ ** 2 * M_ + 4 * N_ * M_ ** 2 + 12 * M_ * N_ + 2 * N_ * M_ + 12 * M_ ** 2 + M_ ** 3 / 3 + 17 * M_
METRIC = "RMCKF 4-feature closed-loop filter-steps/sec"
CPU_TRIALS_PER_CORE = 150          # x 300 frames at ~4e3 frames/s/core: 10-15 s of CPU work per sample
UNIT = "filter-steps/s"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the bench-sized kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/profile_summary.py); None when absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(kernel)
    except (OSError, ValueError):
        return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="dense", choices=["dense", "blockdiag"])
    ap.add_argument("--trials", type=int, default=None, help="trials per GPU per step (weak scaling; default at N = 1: 100 000)")
    ap.add_argument("--total-trials", type=int, default=None,
                    help="trials of the whole job per step, sharded over the ranks (strong scaling; default at N > 1: "
                         "1 000 000 = BASELINE configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-blockdiag-extra", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the trajectory-logging and camera-twist legs")
    return ap.parse_args()


def workload_name(mode, trials, total=None, world=1):
    size = (f"{total} trials per step dealt round-robin to {world} GPUs (BASELINE configs[3])" if total is not None and world > 1
            else f"{trials} trials per GPU per step (BASELINE configs[2])")
    return (f"RMCKF(GMCKF) 4 features / 48 states, closed-loop kinematic UR10 IBVS, alpha-stable noise "
            f"alpha=linspace(1,2,12) beta=0 gamma=1, annealing, kernel_bw=10, T=300 frames, {size}, covariance={mode}")


def plan(a, rank, world):
    """(trials of this rank, first global trial, trial stride, total trials, scaling).  Strong case: the trials are dealt
    round-robin (sweep.shard_interleaved: rank r runs r, r + N, ...) so that every rank has the same mix of noise levels;
    weak case: a contiguous block per rank."""
    if a.trials is None and (world > 1 or a.total_trials is not None):
        total = a.total_trials or 1_000_000
        count = (total - rank + world - 1) // world if total > rank else 0
        return count, rank, world, total, "strong"
    B = a.trials or 100_000
    return B, rank * B, 1, world * B, "weak"


# ------------------------------------------------------------------------- CPU arm (oracle port)
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    trial0, n, method = args
    from oracle import closed_loop as CL, filters as F
    cfg = CL.make_cfg(method=getattr(F, method), annealing=True, noise_seed=NOISE_SEED, pose_seed=POSE_SEED)
    frames = 0
    t0 = time.perf_counter()
    for t in range(trial0, trial0 + n):
        lvl = (t * 12) // 1200 % 12
        cfg["noise_params"] = dict(alpha=SWEEP[lvl], beta=0.0, gamma=1.0, delta=0.0)
        frames += CL.run_trial(cfg, t)["frames"]
    return frames, time.perf_counter() - t0


def cpu_sample(trials_per_core, cores=None, trial0=0, method="GMCKF"):
    """The oracle's as-written NumPy closed loop (experiment.py restated, dense) on all host cores,
    one single-threaded process per core.  Returns (filter-steps/s aggregate, cores, frames, seconds)."""
    cores = cores or os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    ctx = mp.get_context("spawn")
    jobs = [(trial0 + i * trials_per_core, trials_per_core, method) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    work = max(r[1] for r in res)               # slowest worker = the parallel region without pool start-up
    frames = sum(r[0] for r in res)
    return frames / work, cores, frames, wall


def reference_calibration():
    """port / reference ratios measured in the build container (tools/cpu_reference.py -> profiles/cpu_reference.json):
    the unmodified /root/reference Experiment.run() + kinematic fake robot against the oracle port this file times.
    The reference tree does not exist on the GPU box, so the ratio travels as a committed calibration."""
    try:
        with open(os.path.join(ROOT, "profiles", "cpu_reference.json")) as fh:
            return json.load(fh)
    except (OSError, ValueError):
        return None


def cpu_baseline_block(cores=None, reps=3):
    """the port on all host cores, per estimator, best of `reps` bounded samples, with the
    committed port / reference calibration beside it."""
    cal = reference_calibration()
    per = {}
    wall = 0.0
    for method, tpc in (("GMCKF", 60), ("KF", 40), ("MCKF", 24)):
        best = 0.0
        for r in range(reps if method == "GMCKF" else 2):
            v, cores, frames, w = cpu_sample(tpc, cores, trial0=500 * r, method=method)
            best = max(best, v)
            wall += w
        per[method] = best
    block = {"value": per["GMCKF"], "unit": UNIT, "cores": cores, "kind": "port",
             "per_estimator": {("RMCKF(GMCKF)" if k == "GMCKF" else k): v for k, v in per.items()},
             "sample": f"{cores} single-threaded processes x (60 RMCKF / 40 KF / 24 MCKF) closed-loop trials x {T_FRAMES} "
                       f"frames per repetition, best of {reps} (RMCKF) / 2; oracle as-written NumPy restatement of "
                       f"experiment.py; {wall:.0f} s wall in total"}
    if cal:
        bm = cal["by_method"]
        block["port_over_reference"] = bm["GMCKF"]["port_over_reference_allcores"]
        block["port_over_reference_per_estimator"] = {k: v["port_over_reference_allcores"] for k, v in bm.items()}
        block["reference_equivalent"] = per["GMCKF"] / bm["GMCKF"]["port_over_reference_allcores"]
        block["calibration"] = ("port / reference measured in the build container by tools/cpu_reference.py "
                                f"(profiles/r02_cpu_reference.md: unmodified Experiment.run() + kinematic fake robot, "
                                f"{cal['meta']['cores']} cores, best of {cal['meta']['reps']}); the reference tree cannot "
                                "travel to the GPU box")
    return block


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    tpc = CPU_TRIALS_PER_CORE
    for _ in range(a.warmup):
        cpu_sample(1, cores)
    vals, ms = [], []
    for s in range(max(1, a.steps)):
        v, cores, frames, wall = cpu_sample(tpc, cores, trial0=1000 * s)
        vals.append(v)
        ms.append(1e3 * frames / v)
    value = float(np.mean(vals))
    cal = reference_calibration()
    sample = f"{cores} processes x {tpc} closed-loop trials x {T_FRAMES} frames per step (as-written NumPy restatement, OMP_NUM_THREADS=1)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": max(1, a.steps),
        "warmup": a.warmup, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": plan(a, 0, a.gpus)[4],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name("dense", plan(a, 0, a.gpus)[0], plan(a, 0, a.gpus)[3], a.gpus),
                   "trials_per_gpu": plan(a, 0, a.gpus)[0], "frames": T_FRAMES,
                   "sample": f"each step runs {cores * tpc} trials of that workload (bounded CPU sample)",
                   "noise_seed": NOISE_SEED, "pose_seed": POSE_SEED},
        "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                             **({"port_over_reference": cal["by_method"]["GMCKF"]["port_over_reference_allcores"],
                                 "reference_equivalent": value / cal["by_method"]["GMCKF"]["port_over_reference_allcores"]}
                                if cal else {})),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        pw = [float(s[2]) for s in self.samples if s[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples), "reasons": reasons}


# ------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import uvs_b200 as uvs
    L, E = uvs._lib, uvs.engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B, trial0, stride, total_trials, scaling = plan(a, rank, world)   # Philox is keyed by the GLOBAL trial id
    n_levels = len(SWEEP)
    epoch = (total_trials + n_levels - 1) // n_levels

    def make_cfg(mode):
        p = E.filter_params(L.METHOD_GMCKF, 4, mode=L.MODE_DENSE if mode == "dense" else L.MODE_BLOCKDIAG,
                            kernel_bw=10.0, annealing=True)
        return E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=NOISE_SEED, alpha=1.5, beta=0.0,
                                                   gamma=1.0, delta=0.0),
                                 T=T_FRAMES, dt=DT, ibvs_gain=0.2, q_start=Q_START, desired_f=DESIRED_F,
                                 pose_seed=POSE_SEED, sweep=SWEEP, sweep_epoch=epoch, sweep_param=2, trial_stride=stride)

    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")   # > 126 MB L2
    # the per-level statistics block of the job: [n_levels,3,3] sums + [n_levels] failure counts in ONE buffer, so
    # that the only collective of the job is one all-reduce of n_levels * 10 doubles after the last step
    packed = torch.zeros(n_levels * 10, dtype=torch.float64, device="cuda")
    sums_v, fail_v = packed[:n_levels * 9].view(n_levels, 3, 3), packed[n_levels * 9:]

    def step(cfg, out):
        E.closed_loop(cfg, trial0, B, out=out)
        E.reduce_stats(out["stats"], out["status"], trial0, epoch, n_levels, out=(sums_v, fail_v), trial_stride=stride)

    def timed(mode, steps, warmup):
        cfg = make_cfg(mode)
        out = E.closed_loop(cfg, trial0, B)             # allocates the result buffers once
        for _ in range(max(0, warmup - 1)):
            step(cfg, out)
        if world > 1:
            dist.all_reduce(packed)                     # warm the communicator up outside the timed region
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ke = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        packed.zero_()
        e0.record()
        for i in range(steps):                            # no host synchronisation and no collective in here
            flush.fill_(0.0)                              # L2 flush between timed iterations
            ks[i].record()
            E.closed_loop(cfg, trial0, B, out=out)
            ke[i].record()
            E.reduce_stats(out["stats"], out["status"], trial0, epoch, n_levels, out=(sums_v, fail_v), trial_stride=stride)
        if world > 1:
            dist.all_reduce(packed)                       # THE collective: per-level sums of all steps, once
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        kernel_ms = [ks[i].elapsed_time(ke[i]) for i in range(steps)]
        frames = out["frames"].sum(dtype=torch.int64)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(frames)
        return dict(ms_total=float(t.item()), frames=int(frames.item()), kernel_ms=float(np.mean(kernel_ms)),
                    sums=sums_v.cpu().numpy() / steps, n_fail=fail_v.cpu().numpy() / steps,
                    local_frames=int(out["frames"].sum().item()))

    with ClockSampler(local) as clocks:
        main = timed(a.mode, a.steps, a.warmup)
    value = main["frames"] * a.steps / (main["ms_total"] * 1e-3)

    # end-to-end through the host-buffer C-ABI entry point (per step: launch, D2H, sync)
    cfg = make_cfg(a.mode)
    E.closed_loop_host(cfg, trial0, B)       # warm-up at the timed size: page-locked result buffers, stream-ordered pool
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_frames = 0
    for _ in range(a.steps):
        st, status, frames = E.closed_loop_host(cfg, trial0, B)
        e2e_frames += int(frames.sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([e2e_s, float(e2e_frames)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt)
        e2e_value = float(tt[1].item() / tmax[0].item())
    else:
        e2e_value = e2e_frames / e2e_s
    h2d = int(__import__("ctypes").sizeof(L.ClosedLoopCfg))
    d2h = B * (3 * 8 + 4 + 4)

    extra = None
    if not a.no_blockdiag_extra and a.mode == "dense":
        bd = timed("blockdiag", max(1, a.steps), 2)
        extra = {"value": bd["frames"] * max(1, a.steps) / (bd["ms_total"] * 1e-3), "unit": UNIT,
                 "kernel_ms": bd["kernel_ms"],
                 "note": "block-diagonal family (exact for the reference's P0=Q=I, R=I); its roofline uses its OWN "
                         "hardware-counted flops per step, never F_alg"}
        bd_fl = ncu_traffic("blockdiag_executed_flops_per_filter_step")
        if bd_fl:
            extra["flops_per_filter_step_executed"] = bd_fl
            extra["achieved_executed_tflops"] = bd["local_frames"] * bd_fl / (bd["kernel_ms"] * 1e-3) / 1e12

    # trajectory logging (experiment.py:323-330: q, f, noise per frame = 22 doubles at 4 features): the same launch with
    # the [B,T,22] log enabled, its HBM write rate beside the measured copy bandwidth
    logging = twist = None
    if not a.no_extras and a.mode == "dense" and rank == 0:
        Bl = min(B, 50_000)
        cfgl = make_cfg("dense")
        outl = E.closed_loop(cfgl, trial0, Bl, want_log=True)
        torch.cuda.synchronize()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.fill_(0.0)
        l0.record()
        E.closed_loop(cfgl, trial0, Bl, out=outl)
        l1.record()
        torch.cuda.synchronize()
        ms = l0.elapsed_time(l1)
        fr = int(outl["frames"].sum().item())
        log_bytes = fr * (6 + 2 * M_) * 8
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except (OSError, ValueError):
            pass
        logging = {"value": fr / (ms * 1e-3), "unit": UNIT, "trials": Bl, "log_bytes_per_filter_step": (6 + 2 * M_) * 8,
                   "hbm_write_gbs": log_bytes / (ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs"),
                   "note": "dense closed loop with the per-frame trajectory log (q, f, noise) written to HBM"}
        del outl
        # ... and the same log STREAMED to the host (uvs_closed_loop_stream: chunked launches, two device + two page-locked
        # host buffers, the D2H of chunk i under the launch of chunk i + 1) — main.py:152-196 logs every experiment, and
        # this is where PCIe, not HBM, is the limit.  Beside it: the plain pinned D2H rate of this box, measured here.
        try:
            big = torch.empty(1 << 27, dtype=torch.float64, device="cuda")          # 1 GiB
            host = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
            host.copy_(big, non_blocking=True)
            torch.cuda.synchronize()
            l0.record()
            host.copy_(big, non_blocking=True)
            l1.record()
            torch.cuda.synchronize()
            pcie = big.numel() * 8 / (l0.elapsed_time(l1) * 1e-3) / 1e9
            del big, host
            # chunk = 3 whole waves of the dense kernel's persistent grid (quad kernel: 148 SMs x 3 CTAs x 4 filters = 1776
            # per wave): a chunk that ends mid-wave leaves SMs idle at every chunk boundary (measured: 4096 -> 0.84 of the rate)
            CH = 3 * 1776
            E.closed_loop_stream(cfgl, trial0, 2 * CH, chunk=CH)                    # warm-up (allocations, page locking)
            Bs = min(B, 8 * CH)
            si = E.closed_loop_stream(cfgl, trial0, Bs, chunk=CH)
            stream_steps = Bs * T_FRAMES
            logging["stream"] = {"value": stream_steps / si["wall_s"], "unit": UNIT, "trials": Bs, "chunk": CH,
                                 "d2h_gbs_copy_engine": si["d2h_gbs"], "d2h_gbs_sustained": stream_steps * (6 + 2 * M_) * 8 / si["wall_s"] / 1e9,
                                 "pcie_d2h_gbs_measured": pcie,
                                 "note": "dense closed loop with every frame row delivered to page-locked host memory; "
                                         "sustained = log bytes / wall time of the whole pipelined call"}
            import tempfile
            cfg_b = make_cfg("blockdiag")
            si_b = E.closed_loop_stream(cfg_b, trial0, Bs, chunk=CH)
            logging["stream_blockdiag"] = {"value": stream_steps / si_b["wall_s"], "unit": UNIT,
                                           "d2h_gbs_sustained": stream_steps * (6 + 2 * M_) * 8 / si_b["wall_s"] / 1e9,
                                           "note": "block-diagonal family: the kernel produces 190+ GB/s of log, the PCIe copy is the bound"}
            with tempfile.TemporaryDirectory() as td:
                t0 = time.perf_counter()
                st_h, ss_h, fr_h, inf = uvs.sweep.stream_results(cfgl, trial0, trial0 + 480, SWEEP, epoch, chunk=240,
                                                                 csv_path=os.path.join(td, "results.csv"),
                                                                 desired_f=DESIRED_F, dt=DT)
                csv_bytes = os.path.getsize(os.path.join(td, "results.csv"))
            logging["csv"] = {"rows": inf["rows"], "rows_per_s": inf["rows"] / inf["csv_s"], "mb_per_s": csv_bytes / inf["csv_s"] / 1e6,
                              "note": "results.csv rows in the reference schema (41 columns) through the vectorised writer "
                                      "(one DataFrame per chunk, DataFrame.to_csv as main.py:196)"}
        except Exception as e:                                                      # an extra must not sink the bench line
            logging["stream_error"] = repr(e)
        # camera-twist pipeline of the reference's tests/ scripts (kalman_4_features.py configuration, white noise)
        pt = E.filter_params(L.METHOD_KF, 4, mode=L.MODE_DENSE)
        cfgt = E.closed_loop_cfg(pt, E.noise_params(L.NOISE_WHITE, seed=NOISE_SEED, std=1.0), T=440, dt=DT, ibvs_gain=0.5,
                                 q_start=[0.0, -np.pi / 8, np.pi / 2 + np.pi / 8, 0.0, -np.pi / 2, 0.0],
                                 desired_f=DESIRED_F, change_q_start=False)
        tw = E.twist_cfg(x0_mode=L.X0_IMAGE_JACOBIAN, focal_x0=1024.0, err_threshold=0.01)
        Bt = 20_000
        outt = E.twist_loop(cfgt, tw, 0, Bt)
        torch.cuda.synchronize()
        l0.record()
        E.twist_loop(cfgt, tw, 0, Bt, out=outt)
        l1.record()
        torch.cuda.synchronize()
        twist = {"value": int(outt["frames"].sum().item()) / (l0.elapsed_time(l1) * 1e-3), "unit": UNIT, "trials": Bt,
                 "note": "uvs_twist_loop: tests/kalman_4_features.py pipeline (KF on the camera twist, pinv(J_geom) "
                         "joint rates), 440 frames, white noise"}

    peak_tf, _ = E.fp64_peak(4096)
    peak_tf = max(peak_tf, E.fp64_peak(4096)[0])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    steps_per_launch = main["local_frames"]
    achieved = steps_per_launch * F_ALG / (main["kernel_ms"] * 1e-3) / 1e12
    # the dense family's FAST configuration (Q = R = NULL, Joseph) runs the warp-specialised quad kernel (csrc/dense_quad.cuh)
    # unless UVS_DENSE_QUAD=0 asks for the one-filter-per-warp kernel (comparison runs)
    quad = a.mode == "dense" and os.environ.get("UVS_DENSE_QUAD", "1")[:1] != "0"
    kernel_name = "dense_quad_kernel<8, GMCKF>" if quad else f"{a.mode}_closed_loop_kernel<8, GMCKF>"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": main["ms_total"] / a.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a.mode, B, total_trials if scaling == "strong" else None, world),
                   "trials_per_gpu": B, "total_trials": total_trials, "frames": T_FRAMES,
                   "collective": "none inside the step loop; one all-reduce of 120 doubles (per-level sums of all steps) "
                                 "after the last step, inside the timed region" if world > 1 else "none (1 GPU)",
                   "l2": "256 MB buffer written between timed iterations; filter state is shared-memory resident, "
                         "the only HBM traffic is 32 B of results per trial",
                   "noise_seed": NOISE_SEED, "pose_seed": POSE_SEED},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "uvs_closed_loop_host (C ABI: configuration struct in by host pointer, page-locked host result "
                       "buffers out; launch, D2H of the per-trial results and stream synchronisation inside the timed region)"},
        "gpu_launches": 2 * a.steps,      # per step: the closed-loop kernel + the statistics reduction
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved / peak_tf, "traffic": ncu_traffic(kernel_name),
                     "traffic_unit": "bytes per launch (ncu dram read + write); algorithmic: 32 B per trial of results",
                     "kernel": kernel_name, "kernel_ms": main["kernel_ms"],
                     "flops_per_filter_step_algorithmic": F_ALG,
                     "flops_per_filter_step_executed": ncu_traffic("dense_executed_flops_per_filter_step") if a.mode == "dense" else None,
                     "peak_source": "measured in this run: uvs_fp64_peak (register-resident DFMA chains on every SM); "
                                    "MEASURED_PEAKS.json has no FP64 entry; datasheet 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2"},
        "failed_trials": float(main["n_fail"].sum()),
        "itae_mean_per_level": [float(s[2, 1] / max(s[2, 0], 1.0)) for s in main["sums"]],
    }
    ex_fl = line["roofline"]["flops_per_filter_step_executed"]
    if ex_fl:       # hardware-counted (committed ncu capture): executed flops are ~0.6 F_alg because only the lower blocks are updated
        line["roofline"]["achieved_executed"] = steps_per_launch * ex_fl / (main["kernel_ms"] * 1e-3) / 1e12
        line["roofline"]["frac_executed"] = line["roofline"]["achieved_executed"] / peak_tf
    if extra:
        if "achieved_executed_tflops" in extra:
            extra["frac_executed"] = extra["achieved_executed_tflops"] / peak_tf
        line["blockdiag"] = extra
    if logging:
        line["logging"] = logging
    if twist:
        # KF update with dense X, P, K costs the same F_alg; the twist law adds two 6-column pseudo-inverses
        twist["roofline_frac_by_F_alg"] = twist["value"] * F_ALG / 1e12 / peak_tf
        line["twist"] = twist
    if not a.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_block()
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
