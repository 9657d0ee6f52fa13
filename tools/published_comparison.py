#!/usr/bin/env python
"""Closed-loop ITAE of this engine beside the reference's PUBLISHED curves.

    python tools/published_comparison.py [--epoch 4000] [--out profiles/r01_itae_vs_published.md]      (needs a GPU)

The only quantities the reference publishes are the ITAE mean +- std per noise level of results/results{1,2,3}.fig
(decoded from the .fig files; CoppeliaSim in the loop, 100 experiments per level).  This script runs
the same sweep — alpha-stable noise, alpha = linspace(1, 2, 12), 4 features, T = 300, ibvs_gain = 0.2 (config.json) —
for every published series on the kinematic plant (and again with the scenes' first-order servo lag, and with their
rate limits), `epoch` trials per
level, and writes both tables side by side.  The published tables are parsed from BASELINE.md at run time.

What to expect: order-of-magnitude agreement — the plants differ (kinematic UR10 + point features here,
dynamics + rendered discs there) and the published means rest on 100 heavy-tailed samples per level.
"""
import argparse
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402

Q_START = [0.0, 0.0, 1.96349541, 0.0, -1.57079633, 0.0]
DESIRED_F = [149.0, 145.0, 125.0, 121.0, 101.0, 145.0, 125.0, 169.0]
ALPHAS = [float(v) for v in np.linspace(1.0, 2.0, 12)]


def published_tables(path):
    """{figure: {series: [(mean, std)] * 12}} from the markdown tables of BASELINE.md §1."""
    text = open(path, encoding="utf-8").read()
    out = {}
    for fig, body in re.findall(r"### `results/(results\d)\.fig`[^\n]*\n((?:\|[^\n]*\n)+)", text):
        rows = [[c.strip() for c in ln.strip().strip("|").split("|")] for ln in body.strip().splitlines()]
        names = rows[0][1:]
        series = {n: [] for n in names}
        for r in rows[2:]:
            for n, cell in zip(names, r[1:]):
                m = re.match(r"([\d ]+)\s*±\s*([\d ]+)", cell)
                series[n].append((float(m.group(1).replace(" ", "")), float(m.group(2).replace(" ", ""))))
        out[fig] = series
    return out


# published series -> (method, kernel_bw, annealing) ; figure -> random start poses?
_R1 = {"KF": (2, 10.0, False), "RMCKF": (5, 10.0, True), "MCKF": (3, 10.0, True), "IMCC-KF": (4, 10.0, True)}
SERIES = {
    # results1: the .fig does not say what produced it.  Round 1 took it for the fixed start pose (smaller spread at alpha = 2
    # than results2).  The scene decode of round 2 rules that out for config.json's q_start: there the green disc (radius
    # 0.0375 m) is ENTIRELY outside the image, the detector returns NaN and every experiment fails (the scene-plant rows
    # below show 100 % failures).  So results1 also had perturbed start poses; both readings are reported.
    "results1 (read as: fixed start pose)": dict(fig="results1", change_q_start=False, runs=_R1),
    "results1 (read as: random start poses)": dict(fig="results1", change_q_start=True, runs=_R1),
    "results2": dict(fig="results2", change_q_start=True, runs=_R1),
    "results3": dict(fig="results3", change_q_start=True, runs={"KF": (2, 10.0, False), "RMCKF (Annealing)": (5, 10.0, True),
                                                                 "RMCKF (σ = 1)": (5, 1.0, False), "RMCKF (σ = 10)": (5, 10.0, False)}),
}


def run_series(uvs, method, kernel_bw, annealing, change_q_start, epoch, limits):
    L, E = uvs._lib, uvs.engine
    p = E.filter_params(method, 4, mode=L.MODE_AUTO, kernel_bw=kernel_bw, annealing=annealing)
    cfg = E.closed_loop_cfg(p, E.noise_params(L.NOISE_ALPHA_STABLE, seed=123456, alpha=1.5, beta=0.0, gamma=1.0, delta=0.0),
                            T=300, dt=0.05, ibvs_gain=0.2, q_start=Q_START, desired_f=DESIRED_F,
                            change_q_start=change_q_start, pose_seed=12345, sweep=ALPHAS, sweep_epoch=epoch,
                            sweep_param=2, **limits)
    B = epoch * len(ALPHAS)
    out = E.closed_loop(cfg, 0, B)
    itae = out["stats"][:, 2].cpu().numpy().reshape(len(ALPHAS), epoch)
    ok = (out["status"].cpu().numpy() == 0).reshape(len(ALPHAS), epoch)
    res = []
    for lvl in range(len(ALPHAS)):
        v = itae[lvl][ok[lvl]]
        if len(v) == 0:
            res.append((float("nan"), float("nan"), float("nan"), int((~ok[lvl]).sum())))
            continue
        res.append((float(v.mean()), float(v.std()), float(np.median(v)), int((~ok[lvl]).sum())))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epoch", type=int, default=4000)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_itae_vs_published.md"))
    a = ap.parse_args()
    import uvs_b200 as uvs
    pub = published_tables(os.path.join(ROOT, "BASELINE.md"))
    # servo lag: the scenes' joints are position-controlled with P-gain 0.1 at a 5 ms dynamics step, i.e.
    # a joint covers 1 - 0.9^10 = 0.651 of the commanded step within one 50 ms frame
    # scene plant (round 2): the sensor sits ~(0, 0.007, 0.014) m off DH frame 6, the targets are discs of
    # 0.0375 m radius whose visible-region centroid is the feature (clipped to the image; a trial fails only when a disc is
    # entirely out of view), the joints follow a P servo (gain 0.1) at the 5 ms dynamics step under 60 deg/s, 360 deg/s^2
    discs = dict(sensor_offset=(0.0, 0.007, 0.014), disc_radius=0.0375, fov_check=True)
    plants = {"kinematic": {}, "first-order servo lag (servo_gain = 1 - 0.9^10 = 0.651)": dict(servo_gain=1.0 - 0.9 ** 10),
              "servo limits (60 deg/s, 360 deg/s^2)": dict(v_max=1.0472, a_max=6.283),
              "scene plant: offset sensor + finite discs clipped to the image, kinematic joints": dict(discs),
              "scene plant: offset sensor + finite discs + sub-stepped P servo (kp 0.1 x 10, 60 deg/s, 360 deg/s^2)":
                  dict(discs, servo_substeps=10, servo_kp=0.1, v_max=1.0472, a_max=6.283)}
    md = ["# Closed-loop ITAE: this engine vs the reference's published curves", "",
          f"`python tools/published_comparison.py --epoch {a.epoch}` on one B200.  Published = mean ± std of 100 CoppeliaSim "
          "experiments per level, decoded from `results/results{1,2,3}.fig` (BASELINE.md §1).  Ours = mean ± std "
          f"(median) of {a.epoch} kinematic-plant trials per level, alpha-stable noise, 4 features, T = 300; which "
          "config produced each figure is inferred (BASELINE.md; the .fig files do not say): results2 / results3 = random start "
          "poses; results1 is reported under both readings (fixed / random start pose); annealing on the correntropy filters in "
          "all three.  ratio = ours / published mean.  Plants: the ideal kinematic point-feature plant, its servo variants of "
          "round 1, and the scene plant of round 2 (include/uvs_b200.h: sensor offset, finite discs with image clipping, "
          "sub-stepped position servo; constants decoded from scenarios/*.ttt).", ""]
    summary = []
    for label, spec in SERIES.items():
        fig = spec["fig"]
        for plant, limits in plants.items():
            md += [f"## {label} — {plant}", "", "| α | " + " | ".join(f"{n}: published | ours | ratio" for n in spec["runs"]) + " |",
                   "|---|" + "---|---|---|" * len(spec["runs"])]
            cols = {}
            for name, (method, bw, anneal) in spec["runs"].items():
                cols[name] = run_series(uvs, method, bw, anneal, spec["change_q_start"], a.epoch, limits)
            for lvl, alpha in enumerate(ALPHAS):
                cells = []
                for name in spec["runs"]:
                    pm, ps = pub[fig][name][lvl]
                    om, osd, omed, nfail = cols[name][lvl]
                    cells.append(f"{pm:.0f} ± {ps:.0f} | {om:.0f} ± {osd:.0f} ({omed:.0f}) | {om / pm:.2f}" if om == om else
                                 f"{pm:.0f} ± {ps:.0f} | all {nfail} trials FAIL | –")
                md.append(f"| {alpha:.4f} | " + " | ".join(cells) + " |")
            md.append("")
            for name in spec["runs"]:
                r = np.array([cols[name][lvl][0] / pub[fig][name][lvl][0] for lvl in range(len(ALPHAS))])
                # alpha >= 1.18: beyond the two most impulsive levels, whose published means are dominated by a few samples
                nfail = sum(cols[name][lvl][3] for lvl in range(len(ALPHAS)))
                summary.append((label, plant, name, float(np.exp(np.mean(np.log(r[2:])))), float(r[2:].min()), float(r[2:].max()),
                                float(r[0]), float(r[1]), nfail / (a.epoch * len(ALPHAS))))
    md += ["## Summary: ours / published mean ITAE", "",
           "| figure | plant | series | geometric mean, α ≥ 1.18 | min | max | α = 1.00 | α = 1.09 | failed trials |", "|---|---|---|---|---|---|---|---|---|"]
    for s in summary:
        md.append(f"| {s[0]} | {s[1]} | {s[2]} | {s[3]:.2f} | {s[4]:.2f} | {s[5]:.2f} | {s[6]:.2f} | {s[7]:.2f} | {s[8]:.3f} |")
    md.append("")
    with open(a.out, "w", encoding="utf-8") as fh:
        fh.write("\n".join(md) + "\n")
    print("\n".join(md[-(len(summary) + 4):]))


if __name__ == "__main__":
    main()
