"""Oracle (test infrastructure): the per-frame Jacobian estimators, restated.

``step_dense`` follows /root/reference/experiment.py:164-297 statement by
statement (dense matrices, np.linalg.inv / cholesky / pinv exactly where the
reference calls them) and also covers the tests/ variants of SURVEY.md §8a
row 16 (simple-form covariance update, no-skip MCKF).  ``step_blockdiag`` is
the exact block-structured restatement that is
valid when P0, Q are block-diagonal (6x6 blocks) and R is diagonal.

One filter, one frame per call; plain NumPy.
"""
import numpy as np

# experiment.py:6-11
ANALYTICAL, KF, MCKF, IMCCKF, GMCKF = 1, 2, 3, 4, 5

# per-frame flags reported beside the state
FLAG_SKIP_SINGULAR = 1   # experiment.py:233-236 / :292-294
FLAG_SKIP_MAXEPOCH = 2   # experiment.py:246-248
FLAG_CX_SINGULAR = 4     # experiment.py:224 raises out of run() in the reference
FLAG_NOT_SPD = 8         # include/uvs_b200.h: a Cholesky of P or S met a non-positive pivot


def gaussian_kernel(e, bw):
    # utils.py:171-172
    return np.exp(-0.5 * e ** 2 / bw ** 2)


def annealed_bw(kernel_bw, annealing, k, k_max):
    # experiment.py:195-199 (same block at :252-256 and :267-271)
    if annealing:
        return kernel_bw + 100 * (1 - k / k_max)
    return kernel_bw


def regressor(m, h):
    # experiment.py:188   H = kron(I_m, dq^T)
    return np.kron(np.eye(m), np.asarray(h, dtype=float).ravel())


def step_dense(method, X, P, h, Z, Q, R, sigma,
               fpi_threshold=0.1, fpi_epoch_max=1000,
               joseph=True, mckf_skip=True, mckf_recompute_DW=False):
    """One predict + update, as written in the reference.

    X [N], P [N,N], h [6] (zeros on the first frame: experiment.py:183-188),
    Z [m].  Returns X, P, K, kappa, info.  ``kappa`` is the IBVS weight vector
    of experiment.py:306-308 (ones unless GMCKF).
    """
    m = len(Z)
    N = len(X)
    X = np.asarray(X, dtype=float).reshape(N, 1)
    Zc = np.asarray(Z, dtype=float).reshape(m, 1)
    H = regressor(m, h)
    info = {"flags": 0, "epoch": 0}

    # Prediction  (experiment.py:166-167)
    P = P + Q
    K = np.zeros((N, m))
    skip = False
    kappa = np.ones(m)

    if method == KF:
        # experiment.py:192-193
        K = P @ H.T @ np.linalg.inv(H @ P @ H.T + R)
        X = X + K @ (Zc - H @ X)
    elif method == MCKF:
        # experiment.py:203-250
        Bp = np.linalg.cholesky(P)
        Br = np.linalg.cholesky(R)
        B = np.vstack([np.hstack([Bp, np.zeros((N, m))]),
                       np.hstack([np.zeros((m, N)), Br])])
        B_inv = np.linalg.pinv(B)
        difference = np.inf
        epoch = 0
        Xc = X.copy()
        D = B_inv @ np.vstack([X, Zc])
        W = B_inv @ np.vstack([np.eye(N), H])
        while difference > fpi_threshold and epoch < fpi_epoch_max:
            if mckf_recompute_DW:      # tests/mckf_4_features.py:168-169
                D = B_inv @ np.vstack([X, Zc])
                W = B_inv @ np.vstack([np.eye(N), H])
            e = D - W @ Xc
            cx = gaussian_kernel(e[:N, 0], sigma)
            cy = gaussian_kernel(e[N:, 0], sigma)
            if np.any(cx == 0.0):
                # np.linalg.inv(Cx) at experiment.py:224 is outside the try:
                # the reference would raise LinAlgError out of run().  Flag it
                # and treat the frame as a skipped correction.
                info["flags"] |= FLAG_CX_SINGULAR
                skip = True
                break
            P_hat = Bp @ np.diag(1.0 / cx) @ Bp.T
            if np.any(cy == 0.0):
                if mckf_skip:
                    # experiment.py:233-236
                    info["flags"] |= FLAG_SKIP_SINGULAR
                    skip = True
                    break
                raise np.linalg.LinAlgError("Singular matrix")
            R_hat = Br @ np.diag(1.0 / cy) @ Br.T
            K = P_hat @ H.T @ np.linalg.inv(H @ P_hat @ H.T + R_hat)
            Xc_old = Xc.copy()
            Xc = X + K @ (Zc - H @ X)
            with np.errstate(divide="ignore", invalid="ignore"):
                difference = np.linalg.norm(Xc - Xc_old) / np.linalg.norm(Xc_old)
            epoch += 1
            if epoch == fpi_epoch_max and mckf_skip:
                # experiment.py:246-248
                info["flags"] |= FLAG_SKIP_MAXEPOCH
                skip = True
        info["epoch"] = epoch
        if not skip:
            X = Xc
    elif method == IMCCKF:
        # experiment.py:258-265
        innov = Zc - H @ X
        norm_innov = np.sqrt(innov.T @ np.linalg.inv(R) @ innov)
        c = gaussian_kernel(norm_innov, sigma)
        R_e = c * H @ P @ H.T + R
        K = c * P @ H.T @ np.linalg.inv(R_e)
        X = X + K @ innov
    elif method == GMCKF:
        # experiment.py:121-122, 274-291
        Br = np.linalg.cholesky(R)
        Br_inv = np.linalg.inv(Br)
        e = Br_inv @ Zc - Br_inv @ H @ X
        cy = gaussian_kernel(e[:, 0], sigma)
        Cy_inv = np.linalg.inv(np.diag(cy) + 0.001 ** 2 * np.eye(m))
        R_hat = Br @ Cy_inv @ Br.T
        S = H @ P @ H.T + R_hat
        K = P @ H.T @ np.linalg.inv(S)
        X = X + K @ (Zc - H @ X)
        kappa = cy.copy()          # experiment.py:307-308
    else:
        raise ValueError("method")

    if not skip:
        I_KH = np.eye(N) - K @ H
        if joseph:
            P = I_KH @ P @ I_KH.T + K @ R @ K.T      # experiment.py:297
        else:
            P = I_KH @ P                              # tests/kalman_1_feature.py:126
    return X.ravel(), P, K, kappa, info


def step_blockdiag(method, X, Pb, h, Z, qd, rd, sigma,
                   fpi_threshold=0.1, fpi_epoch_max=1000,
                   joseph=True, mckf_skip=True):
    """Exact block-structured restatement.

    Pb [m,6,6] diagonal blocks of P, qd [m,6,6] diagonal blocks of Q, rd [m]
    diagonal of R.  Valid when P0,Q are block-diagonal and R diagonal, which
    the reference's hard-coded identities always are (experiment.py:73-76).
    """
    m = len(Z)
    X = np.asarray(X, dtype=float).reshape(m, 6).copy()
    h = np.asarray(h, dtype=float)
    Pb = Pb + qd
    v = Pb @ h                      # [m,6]   (P^- H^T, one column per block)
    s0 = v @ h                      # [m]     diag(H P^- H^T)
    nu = Z - X @ h                  # [m]
    info = {"flags": 0, "epoch": 0}
    kappa = np.ones(m)
    skip = False
    K = np.zeros((m, 6))
    if method == KF:
        K = v / (s0 + rd)[:, None]
        Xn = X + K * nu[:, None]
    elif method == GMCKF:
        e = nu / np.sqrt(rd)
        cy = gaussian_kernel(e, sigma)
        K = v / (s0 + rd / (cy + 1e-6))[:, None]
        Xn = X + K * nu[:, None]
        kappa = cy
    elif method == IMCCKF:
        c = gaussian_kernel(np.sqrt(np.sum(nu ** 2 / rd)), sigma)
        K = c * v / (c * s0 + rd)[:, None]
        Xn = X + K * nu[:, None]
    elif method == MCKF:
        L = np.linalg.cholesky(Pb)               # [m,6,6]
        sr = np.sqrt(rd)
        Xc = X.copy()
        difference = np.inf
        epoch = 0
        while difference > fpi_threshold and epoch < fpi_epoch_max:
            ex = np.stack([np.linalg.solve(L[i], X[i] - Xc[i]) for i in range(m)])
            ey = (Z - Xc @ h) / sr
            cx = gaussian_kernel(ex, sigma)
            cy = gaussian_kernel(ey, sigma)
            if np.any(cx == 0.0):
                info["flags"] |= FLAG_CX_SINGULAR
                skip = True
                break
            if np.any(cy == 0.0):
                if mckf_skip:
                    info["flags"] |= FLAG_SKIP_SINGULAR
                    skip = True
                    break
                raise np.linalg.LinAlgError("Singular matrix")
            Ph = np.einsum("mij,mj,mkj->mik", L, 1.0 / cx, L)
            vh = Ph @ h
            K = vh / (vh @ h + rd / cy)[:, None]
            Xc_old = Xc.copy()
            Xc = X + K * nu[:, None]
            with np.errstate(divide="ignore", invalid="ignore"):
                difference = np.linalg.norm(Xc - Xc_old) / np.linalg.norm(Xc_old)
            epoch += 1
            if epoch == fpi_epoch_max and mckf_skip:
                info["flags"] |= FLAG_SKIP_MAXEPOCH
                skip = True
        info["epoch"] = epoch
        Xn = Xc
    else:
        raise ValueError("method")
    if not skip:
        X = Xn
        if joseph:
            sJ = s0 + rd
            Pb = (Pb - K[:, :, None] * v[:, None, :] - v[:, :, None] * K[:, None, :]
                  + sJ[:, None, None] * K[:, :, None] * K[:, None, :])
        else:
            Pb = Pb - K[:, :, None] * v[:, None, :]
    return X.reshape(-1), Pb, K, kappa, info


def blocks_of(P, m):
    """[N,N] -> [m,6,6] diagonal blocks."""
    return np.stack([P[6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(m)])


def off_block_max(P, m):
    Pc = np.array(P, dtype=float, copy=True)
    for i in range(m):
        Pc[6 * i:6 * i + 6, 6 * i:6 * i + 6] = 0.0
    return float(np.max(np.abs(Pc))) if Pc.size else 0.0
