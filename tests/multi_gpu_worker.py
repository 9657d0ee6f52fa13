"""Worker of tests/test_gpu_multi.py (launched under torchrun): runs this rank's shard of a fixed closed-loop job and
lets rank 0 store the gathered per-trial table and the all-reduced per-level sums.

    torchrun --nproc-per-node G tests/multi_gpu_worker.py <out.npz> <n_trials> <mode> <contiguous|interleaved>

One rank per GPU over NCCL when the box has at least G GPUs; otherwise the ranks share the GPUs round-robin and the
reduction runs over gloo on CPU tensors (NCCL refuses two ranks on one device) — the sharding and the kernels are the
same either way.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path, n_trials, mode, shard = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    import uvs_b200 as uvs
    L, E, S = uvs._lib, uvs.engine, uvs.sweep
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = torch.cuda.device_count()
    nccl = ndev >= world
    torch.cuda.set_device(rank % ndev)
    dist.init_process_group("nccl" if nccl else "gloo")
    if shard == "interleaved":                       # rank r runs trials r, r + W, ... (bench.py's strong-scaling split)
        first, stride, count = S.shard_interleaved(n_trials, rank, world)
    else:                                            # contiguous ranges (sweep.run_sweep)
        lo, hi = S.shard_range(n_trials, rank, world)
        first, stride, count = lo, 1, hi - lo
    sweep = [float(v) for v in np.linspace(1.0, 2.0, 12)]
    epoch = (n_trials + 11) // 12
    p = E.filter_params(L.METHOD_GMCKF, 4, mode=L.MODE_DENSE if mode == "dense" else L.MODE_BLOCKDIAG, kernel_bw=10.0,
                        annealing=True)
    cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=123456, alpha=1.5), T=300, dt=0.05, ibvs_gain=0.2,
                            q_start=[0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0],
                            desired_f=[149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0], pose_seed=12345,
                            sweep=sweep, sweep_epoch=epoch, sweep_param=2, trial_stride=stride)
    out = E.closed_loop(cfg, first, count)
    sums, n_fail = E.reduce_stats(out["stats"], out["status"], first, epoch, 12, trial_stride=stride)
    stats, status = out["stats"], out["status"]
    if not nccl:
        sums, n_fail, stats, status = sums.cpu(), n_fail.cpu(), stats.cpu(), status.cpu()
    S.merge_level_sums(sums, n_fail)                       # the job's one all-reduce
    stats, status = S.gather_trials(stats, status)       # rank-major concatenation
    if shard == "interleaved":                           # back to global trial order
        ids = torch.cat([torch.arange(S.shard_interleaved(n_trials, r, world)[2]) * world + r for r in range(world)])
        order = torch.argsort(ids).to(stats.device)
        stats, status = stats[order], status[order]
    if rank == 0:
        np.savez(out_path, stats=stats.cpu().numpy(), status=status.cpu().numpy(), sums=sums.cpu().numpy(),
                 n_fail=n_fail.cpu().numpy(), backend=np.array("nccl" if nccl else "gloo"), world=world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
