"""BASELINE configs[3] on hardware: the per-trial results and the per-level sums of a fixed closed-loop job do not depend
on the number of ranks it is sharded over (contiguous trial ranges, /root/reference/main.py:121-139 ordering; Philox
counters keyed by the global trial id).  Launches torchrun for G in {1, 2, 4, 8}: one rank per GPU over
NCCL when the box has G GPUs, otherwise the ranks share the GPU(s) and reduce over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_gpu_worker.py")


def _run(world, out, n_trials, mode, shard, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), WORKER, str(out), str(n_trials), mode, shard]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    with np.load(out) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("mode,n_trials,shard", [("blockdiag", 96_000, "contiguous"), ("dense", 12_000, "interleaved")])
def test_results_do_not_depend_on_the_number_of_gpus(tmp_path, mode, n_trials, shard):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    ndev = torch.cuda.device_count()
    worlds = [1, 2] + [g for g in (4, 8) if g <= ndev]
    res = {}
    for i, g in enumerate(worlds):
        res[g] = _run(g, tmp_path / f"g{g}.npz", n_trials, mode, shard, 29610 + 7 * i + (0 if mode == "dense" else 40))
    base = res[1]
    assert base["stats"].shape == (n_trials, 3) and (base["status"] == 0).mean() > 0.5
    for g in worlds[1:]:
        assert np.array_equal(res[g]["stats"], base["stats"]), (mode, g)          # bit-identical per trial
        assert np.array_equal(res[g]["status"], base["status"]), (mode, g)
        assert np.array_equal(res[g]["n_fail"], base["n_fail"])
        assert np.array_equal(res[g]["sums"][:, :, 0], base["sums"][:, :, 0])      # counts
        rel = np.abs(res[g]["sums"] - base["sums"]) / np.maximum(np.abs(base["sums"]), 1e-300)
        assert rel.max() <= 1e-12, (mode, g, rel.max())                          # sums: atomics reorder the additions
        if g <= ndev:
            assert str(res[g]["backend"]) == "nccl"
