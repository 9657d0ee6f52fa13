"""GPU diagnostic: per-frame error profile of uvs_twist_loop against the ref_twist goldens."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import load_golden, rel_err
from test_oracle_vs_golden import twist_cfg_from_golden
import test_gpu_twist as TG
import uvs_b200 as uvs
L, E = uvs._lib, uvs.engine
for name in sys.argv[1:]:
    g = load_golden(f"ref_twist_{name}.npz")
    ocfg = twist_cfg_from_golden(g)
    nf, T, k = int(g["nf"]), int(g["T"]), int(g["frames"]); m = 2 * nf
    cfg, tw = TG.gpu_cfg(uvs, ocfg)
    noise = np.zeros((1, T, m)); noise[:, :k] = g["noise_log"]
    X0 = E.dev(g["x0"][None]) if "x0" in g else None
    out = E.twist_loop(cfg, tw, 0, 1, want_log=True, noise_in=E.dev(noise), X0_in=X0)
    log = out["log"].cpu().numpy()[0]
    eq = np.array([rel_err(log[i, :6], g["q_log"][i]) for i in range(k)])
    dq = np.abs(np.diff(g["q_log"], axis=0) / 0.05).max(axis=1)
    print(name, "flags", int(out["flags"][0]), "frames", int(out["frames"][0]))
    for i in list(range(0, min(k, 60))) + list(range(60, k, 20)):
        print(f"  k={i:4d} eq={eq[i]:.2e} |dq|={dq[min(i, k-2)]:.2e} |noise|={np.abs(g['noise_log'][i]).max():.1f}")
