import sys, numpy as np
sys.path.insert(0, '.')
import uvs_b200 as uvs
from oracle import closed_loop as CL, noise_model as NM
L, E = uvs._lib, uvs.engine
np.set_printoptions(linewidth=200, precision=6)
ocfg = CL.make_cfg(method=2, annealing=False, noise_type=NM.GAUSSIAN_MIXTURE, noise_params=dict(std=1.0, mean=50.0, rho=0.1), hold_cnt=10, noise_seed=4321, pose_seed=99)
B, trial0 = 6, 1000
ref = [CL.run_trial(ocfg, trial0 + b) for b in range(B)]
for mode in (L.MODE_DENSE, L.MODE_BLOCKDIAG):
    p = E.filter_params(2, 4, mode=mode)
    cfg = E.closed_loop_cfg(p, E.noise_params(2, seed=4321, hold_cnt=10, std=1.0, mean=50.0, rho=0.1), T=300, dt=0.05, ibvs_gain=0.2, q_start=ocfg["q_start"], desired_f=ocfg["desired_f"], pose_seed=99)
    out = E.closed_loop(cfg, trial0, B, want_log=True)
    print("mode", mode, "status", out["status"].tolist(), "frames", out["frames"].tolist(), "flags", out["flags"].tolist())
    print("ref status", [r["status"] for r in ref], [r["frames"] for r in ref])
    log = out["log"].cpu().numpy()
    for b in range(B):
        k = int(out["frames"][b])
        n = min(k, ref[b]["frames"])
        dn = np.abs(log[b, :n, 14:22] - ref[b]["noise_log"][:n]).max(axis=1)
        df = np.abs(log[b, :n, 6:14] - ref[b]["f_log"][:n]).max(axis=1)
        dqq = np.abs(log[b, :n, 0:6] - ref[b]["q_log"][:n]).max(axis=1)
        bad = np.nonzero((df > 1e-6) | (dn > 1e-9))[0]
        print(" trial", b, "frames", k, "first bad frame", bad[:3], "max dnoise", dn.max() if n else None, "max df", df.max() if n else None, "max dq", dqq.max() if n else None)
        if len(bad):
            kk = bad[0]
            print("   gpu  f", log[b, kk, 6:14]); print("   ref  f", ref[b]["f_log"][kk]); print("   gpu nz", log[b, kk, 14:22]); print("   ref nz", ref[b]["noise_log"][kk])
            print("   gpu q", log[b, kk, :6], " ref q", ref[b]["q_log"][kk])
        if k < 300 and k > 0:
            print("   last frames gpu q", log[b, max(0,k-2):k, :6]); print("   f", log[b, max(0,k-2):k, 6:14])
