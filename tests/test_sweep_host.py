"""Host-side logic of the sharded sweep: sharding, the per-level reduction and its
all-reduce (gloo, world_size 2, CPU tensors), level statistics vs the oracle's plot_errorbar.m restatement,
config handling and the results.csv schema of main.py:152-194.  No CUDA calls."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import uvs_b200                                    # noqa: E402
from oracle import stats as ostats                  # noqa: E402

sweep = uvs_b200.sweep

# main.py:152-194, typed out independently of sweep.csv_columns()
REFERENCE_COLUMNS = ("experiment_id status rho t q_1 q_2 q_3 q_4 q_5 q_6 camera_x camera_y camera_z camera_roll "
                     "camera_pitch camera_yaw f_1 f_2 f_3 f_4 f_5 f_6 f_7 f_8 desired_f_1 desired_f_2 desired_f_3 "
                     "desired_f_4 desired_f_5 desired_f_6 desired_f_7 desired_f_8 noise_1 noise_2 noise_3 noise_4 "
                     "noise_5 noise_6 noise_7 noise_8 kernel_bw").split()


def _fake_trials(n, seed=0):
    rng = np.random.default_rng(seed)
    stats = rng.lognormal(8.0, 1.0, size=(n, 3))
    status = (rng.uniform(size=n) < 0.1).astype(np.int32)
    return stats, status


def test_csv_schema_is_the_references():
    assert sweep.csv_columns(8) == REFERENCE_COLUMNS
    assert len(sweep.csv_columns(6)) == 4 + 6 + 6 + 3 * 6 + 1


@pytest.mark.parametrize("n,world", [(1200, 1), (1200, 8), (1201, 8), (7, 8), (0, 2), (1_000_000, 4)])
def test_shard_range_partitions_the_queue(n, world):
    parts = [sweep.shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
    sizes = [hi - lo for lo, hi in parts]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        sweep.shard_range(n, world, world)


def test_level_sums_and_summary_match_oracle_statistics():
    epoch, n_levels = 37, 12
    stats, status = _fake_trials(epoch * n_levels - 5, seed=3)         # ragged last level
    sums, n_fail = sweep.level_sums_host(stats, status, 0, epoch, n_levels)
    summ = sweep.level_summary(sums, n_fail, levels=np.linspace(1, 2, 12))
    quant = sweep.level_quantiles(stats, status, epoch, n_levels)
    for lv in range(n_levels):
        sl = slice(lv * epoch, min((lv + 1) * epoch, len(stats)))
        ref = ostats.config_stats(stats[sl], status[sl])
        assert summ[lv]["n_ok"] == ref["n_ok"] and summ[lv]["n_fail"] == ref["n_fail"]
        for j, name in enumerate(("ise", "iae", "itae")):
            assert np.isclose(summ[lv][name + "_mean"], ref[name + "_mean"], rtol=1e-12)
            assert np.isclose(summ[lv][name + "_std"], ref[name + "_std"], rtol=1e-9)
            assert np.isclose(quant[lv, j, 1], ref[name + "_median"], rtol=1e-12)
    # trial offset: a shard that starts mid-level lands in the right levels
    s2, f2 = sweep.level_sums_host(stats[50:], status[50:], 50, epoch, n_levels)
    s1, f1 = sweep.level_sums_host(stats[:50], status[:50], 0, epoch, n_levels)
    assert torch.allclose(s1 + s2, sums, rtol=1e-13) and torch.equal(f1 + f2, n_fail)


def test_level_summary_edge_cases():
    sums = torch.zeros((2, 3, 3), dtype=torch.float64)
    sums[1, :, 0], sums[1, :, 1], sums[1, :, 2] = 1.0, 5.0, 25.0
    out = sweep.level_summary(sums, torch.tensor([4.0, 0.0]))
    assert out[0]["n_ok"] == 0 and out[0]["n_fail"] == 4 and np.isnan(out[0]["itae_mean"])
    assert out[1]["itae_mean"] == 5.0 and np.isnan(out[1]["itae_std"])            # one sample: MATLAB std -> NaN here


def _gloo_worker(rank, world, port, n_trials, epoch, n_levels, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stats, status = _fake_trials(n_trials, seed=11)               # every rank can build the whole queue
        lo, hi = sweep.shard_range(n_trials, rank, world)
        sums, n_fail = sweep.level_sums_host(stats[lo:hi], status[lo:hi], lo, epoch, n_levels)
        sweep.merge_level_sums(sums, n_fail)
        gs, gst = sweep.gather_trials(torch.as_tensor(stats[lo:hi]), torch.as_tensor(status[lo:hi]))
        q.put((rank, sums.numpy(), n_fail.numpy(), gs.numpy(), gst.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_trials", [1200, 1201])
def test_two_rank_reduction_equals_single_process(n_trials):
    """world_size 2 over gloo: sharded partial sums + all-reduce == the unsharded reduction, on both ranks;
    the gathered per-trial table is the original queue in trial order (ragged shards included)."""
    epoch, n_levels, world = 100, 12, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n_trials, epoch, n_levels, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    stats, status = _fake_trials(n_trials, seed=11)
    ref_sums, ref_fail = sweep.level_sums_host(stats, status, 0, epoch, n_levels)
    for rank, sums, n_fail, gs, gst in got:
        assert np.allclose(sums, ref_sums.numpy(), rtol=1e-13)
        assert np.array_equal(n_fail, ref_fail.numpy())
        assert np.array_equal(gs, stats) and np.array_equal(gst, status)


# the shipped experiment description (config.json keys of main.py:16-97; values of SURVEY.md §8d config 3)
BASE_CONFIG = {
    "log_level": "INFO",
    "experiments": {"dt": 0.05, "t_max": 15, "epoch": 100, "ibvs_gain": 0.2,
                    "q_start": [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0],
                    "desired_f": [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0],
                    "visualization": False, "change_q_start": True, "seed": 12345},
    "estimator": {"method": "MCKF", "estimator_params": {"initial_guess": True, "kernel_bw": 10, "fpi_threshold": 0.1,
                                                         "fpi_epoch_max": 1000, "annealing": False}},
    "noise": {"type": "ALPHA_STABLE", "noise_params": {"beta": 0, "gamma": 1, "delta": 0}, "hold": False,
              "hold_time": 0.5, "seed": 123456},
}


def _config(**over):
    cfg = json.loads(json.dumps(BASE_CONFIG))
    for path, v in over.items():
        d = cfg
        keys = path.split(".")
        for k in keys[:-1]:
            d = d[k]
        if v is None:
            d.pop(keys[-1])
        else:
            d[keys[-1]] = v
    return cfg


def test_load_config_follows_main_py():
    run = sweep.load_config(_config())
    assert run["method"] == uvs_b200.Method.MCKF and run["noise_type"] == uvs_b200.NoiseType.ALPHA_STABLE
    assert run["sweep_param"] == "alpha" and np.allclose(run["rho_list"], np.linspace(1, 2, 12))     # main.py:104-106
    assert run["epoch"] == 100 and run["dt"] == 0.05 and run["t_max"] == 15 and run["noise_seed"] == 123456
    run = sweep.load_config(_config(**{"noise.type": "GAUSSIAN_MIXTURE", "estimator.method": "GMCKF"}))
    assert run["sweep_param"] == "rho" and np.allclose(run["rho_list"], np.linspace(0, 0.2, 12))
    assert run["method"] == uvs_b200.Method.RMCKF                       # alias of GMCKF
    for bad in ({"estimator.method": "EKF"}, {"noise.type": "PINK"}, {"experiments.epoch": None},
                {"noise.hold_time": None}):
        with pytest.raises(SystemExit):                                  # main.py:71-73,88-90,99-101
            sweep.load_config(_config(**bad))


def test_load_config_optional_plant_block():
    """Additive: "plant": {servo_gain, v_max, a_max} reaches the estimator kwargs; a config without it is the kinematic plant."""
    cfg = _config()
    assert not {"servo_gain", "v_max", "a_max"} & set(sweep.load_config(cfg)["method_params"])
    cfg["plant"] = {"servo_gain": 0.651, "v_max": 1.0472}
    mp_ = sweep.load_config(cfg)["method_params"]
    assert mp_["servo_gain"] == 0.651 and mp_["v_max"] == 1.0472 and "a_max" not in mp_
    cfg["plant"] = {"friction": 1.0}
    with pytest.raises(SystemExit):
        sweep.load_config(cfg)


def test_run_sweep_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(uvs_b200._lib.UvsError):
        sweep.run_sweep(_config(**{"estimator.method": "GMCKF"}), epoch=1)


def test_published_tables_parse_and_servo_rate():
    """tools/published_comparison.py reads the reference's published ITAE curves from BASELINE.md (3 figures x 4 series x
    12 noise levels); oracle.closed_loop.servo_rate is the plant model both oracles and the kernels share."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import published_comparison as pc
    from oracle import closed_loop as CL
    pub = pc.published_tables(os.path.join(ROOT, "BASELINE.md"))
    assert sorted(pub) == ["results1", "results2", "results3"]
    for spec in pc.SERIES.values():              # every reported series has its published column
        series = pub[spec["fig"]]
        assert set(series) == set(spec["runs"]) and all(len(v) == 12 for v in series.values())
    assert pub["results2"]["RMCKF"][0] == (233102.0, 355742.0) and pub["results3"]["RMCKF (σ = 1)"][11][0] == 238561.0
    cfg = dict(dt=0.05, v_max=1.0, a_max=4.0, servo_gain=0.5)
    v = CL.servo_rate(cfg, np.array([4.0, -0.1, 0.3]), np.array([0.0, 0.0, 0.2]))
    assert np.allclose(v, [0.2, -0.05, 0.15])          # lag 0.5, then |dv| <= a_max dt = 0.2, then |v| <= 1
    assert np.allclose(CL.servo_rate(dict(dt=0.05), np.array([3.0]), np.array([0.0])), [3.0])      # no limits: kinematic


def test_shipped_config_is_the_reference_configuration():
    """`python main.py` runs bare, as /root/reference/main.py:16 does: the repository ships a config.json with the
    reference's values (SURVEY.md section 5 lists every key)."""
    with open(os.path.join(ROOT, "config.json")) as fh:
        shipped = json.load(fh)
    assert shipped == BASE_CONFIG
    run = sweep.load_config(os.path.join(ROOT, "config.json"))
    assert run["epoch"] == 100 and run["method"] == uvs_b200.Method.MCKF


def test_bench_plan_shards_like_the_driver():
    """bench.py: N = 1 -> configs[2] (100 000 trials, weak); N > 1 -> configs[3] (1 000 000 trials sharded, strong) with
    the contiguous ranges of sweep.shard_range; --trials forces a fixed count per GPU."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a = argparse.Namespace(trials=None, total_trials=None)
    assert bench.plan(a, 0, 1) == (100_000, 0, 1, 100_000, "weak")
    for world in (2, 4, 8, 3):
        got = [bench.plan(a, r, world) for r in range(world)]
        assert all(g[4] == "strong" and g[3] == 1_000_000 and g[2] == world for g in got)
        assert [(g[1], g[2], g[0]) for g in got] == [sweep.shard_interleaved(1_000_000, r, world) for r in range(world)]
        ids = np.concatenate([g[1] + g[2] * np.arange(g[0]) for g in got])
        assert len(ids) == 1_000_000 and len(np.unique(ids)) == 1_000_000 and ids.max() == 999_999     # a partition
    a = argparse.Namespace(trials=5000, total_trials=None)
    assert bench.plan(a, 3, 4) == (5000, 15000, 1, 20000, "weak")
    a = argparse.Namespace(trials=None, total_trials=96_000)
    assert bench.plan(a, 0, 1) == (96_000, 0, 1, 96_000, "strong")
