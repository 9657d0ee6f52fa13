"""GPU diagnostic: per-call wall time of uvs_closed_loop_host at the bench size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import uvs_b200 as uvs
L, E = uvs._lib, uvs.engine
B = 100000
p = E.filter_params(L.METHOD_GMCKF, 4, mode=L.MODE_DENSE, kernel_bw=10.0, annealing=True)
cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=123456, alpha=1.5), T=300, dt=0.05, ibvs_gain=0.2,
                        q_start=bench.Q_START, desired_f=bench.DESIRED_F, pose_seed=12345, sweep=bench.SWEEP,
                        sweep_epoch=B // 12 + 1, sweep_param=2)
for i in range(8):
    t0 = time.perf_counter()
    st, status, frames = E.closed_loop_host(cfg, 0, B)
    t1 = time.perf_counter()
    print(f"call {i}: {1e3 * (t1 - t0):.1f} ms  frames {int(frames.sum())}", flush=True)
out = E.closed_loop(cfg, 0, B)
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); E.closed_loop(cfg, 0, B, out=out); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"device call {i}: {1e3 * (t1 - t0):.1f} ms")
