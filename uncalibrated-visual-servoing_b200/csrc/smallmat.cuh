// smallmat.cuh — register-resident small dense solves (fully unrolled, one thread).
// Used redundantly by every lane of a warp where a filter needs an m x m (m <= 8)
// factorisation: cheaper than a shuffle/smem-cooperative factorisation at this size
// because it needs no synchronisation and its ~8 dependent rsqrt steps overlap with
// other warps' DFMA streams.
#pragma once
#include "common.cuh"

namespace uvs {

// hand-over from the Cholesky path of the pseudo-inverse to the rank-revealing one: pivot <= kPinvPivotTol * trace(Gram)
constexpr double kPinvPivotTol = 1e-11;
// ... and a posteriori: the refinement step's correction measures the error of the first normal-equations solve
// (~ eps cond^2); the refined result is good to about its square.  Corrections above 3e-5 of the solution (cond >~ 4e5)
// go to the rank-revealing path too.  (A pivot test alone underestimates cond(J) whenever the weakest singular vector
// has little weight on the last-eliminated coordinate: measured 2e-6 errors at cond = 1e7.)
constexpr double kPinvRefineTol2 = 3e-5 * 3e-5;

// packed lower-triangular index
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// 1/sqrt(d) for finite d > 0: MUFU.RSQ64H seed + two Newton steps (~1 ulp), without the special-case
// handling of rsqrt() (callers test the pivot themselves).  ~8 instructions instead of ~25.
UVS_DEV double fast_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double hd = 0.5 * d;
    double t = fma(-hd * y, y, 0.5);
    y = fma(y, t, y);
    t = fma(-hd * y, y, 0.5);
    y = fma(y, t, y);
    return y;
}

// NOTE on loops: triangular loops are written with a constant trip count and a predicate
// (`for k < M) if (k < j)`) — nvcc does not fully unroll loops whose bound is an outer induction variable,
// and a partially unrolled loop puts the packed matrix in local memory.
// In-place Cholesky of a packed lower SPD matrix held in registers.
// On exit L holds the factor with its DIAGONAL REPLACED BY ITS RECIPROCAL.
// Returns false when a pivot is not positive/finite (np.linalg.cholesky would raise).
template <int M>
UVS_DEV bool chol_packed(double (&L)[M * (M + 1) / 2]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double d = L[tri(j, j)];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k < j) d -= L[tri(j, k)] * L[tri(j, k)];
        ok = ok && (d > 0.0) && isfinite(d);
        const double inv = fast_rsqrt(d);
        L[tri(j, j)] = inv;
#pragma unroll
        for (int i = 0; i < M; ++i) if (i > j) {
            double s = L[tri(i, j)];
#pragma unroll
            for (int k = 0; k < M; ++k) if (k < j) s -= L[tri(i, k)] * L[tri(j, k)];
            L[tri(i, j)] = s * inv;
        }
    }
    return ok;
}

// x <- (L L^T)^-1 x   with L from chol_packed (reciprocal diagonal)
template <int M>
UVS_DEV void chol_solve(const double (&L)[M * (M + 1) / 2], double (&x)[M]) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double s = x[j];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k < j) s -= L[tri(j, k)] * x[k];
        x[j] = s * L[tri(j, j)];
    }
#pragma unroll
    for (int j = M - 1; j >= 0; --j) {
        double s = x[j];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k > j) s -= L[tri(k, j)] * x[k];
        x[j] = s * L[tri(j, j)];
    }
}

// 1/d for finite normal d: MUFU.RCP64H seed + two Newton steps (~1 ulp), no special-case handling (callers test the pivot).
UVS_DEV double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    t = fma(-d, y, 1.0);
    y = fma(y, t, y);
    return y;
}

// In-place square-root-free factorisation A = L D L^T of a packed lower SPD matrix held in registers.  On exit the strictly
// lower entries hold the UNIT lower factor and the diagonal holds 1 / d_j.  Same pivots d_j as the Cholesky factorisation
// (d_j = l_jj^2), so the same positivity test; returns false when a pivot is not positive / finite.
// Why not Cholesky: the factorisations of S (M x M) and of the Gram matrix (6 x 6) are redundant per lane and sit on the
// frame's critical path.  Per column the dependent chain is pivot -> reciprocal (MUFU + 4 FMA instead of rsqrt's MUFU + 6)
// -> scale, and every substitution step of the solves loses its multiplication by the reciprocal diagonal (unit
// diagonal: one FMA per step instead of FMA + MUL) — ~600 dependent cycles per filter-step (tools/phase_clocks.py).
// The last column's UNSCALED entries u_i = l_ij d_j are kept so that the next pivot's last term does not wait for a
// scale-then-rescale round trip.
template <int M>
UVS_DEV bool ldl_packed(double (&L)[M * (M + 1) / 2]) {
    bool ok = true;
    double dd[M], u[M];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double w[M];                                   // row j of L D: w_k = l_jk d_k
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < j) w[k] = (k == j - 1) ? u[j] : L[tri(j, k)] * dd[k];
        double d = L[tri(j, j)];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k < j) d -= L[tri(j, k)] * w[k];
        ok = ok && (d > 0.0) && isfinite(d);
        dd[j] = d;
        const double inv = fast_rcp(d);
        L[tri(j, j)] = inv;
#pragma unroll
        for (int i = 0; i < M; ++i) if (i > j) {
            double s = L[tri(i, j)];
#pragma unroll
            for (int k = 0; k < M; ++k) if (k < j) s -= L[tri(i, k)] * w[k];
            u[i] = s;
            L[tri(i, j)] = s * inv;
        }
    }
    return ok;
}

// x <- (L D L^T)^-1 x   with L from ldl_packed (unit lower, reciprocal pivots on the diagonal)
template <int M>
UVS_DEV void ldl_solve(const double (&L)[M * (M + 1) / 2], double (&x)[M]) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double s = x[j];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k < j) s -= L[tri(j, k)] * x[k];
        x[j] = s;
    }
#pragma unroll
    for (int j = 0; j < M; ++j) x[j] *= L[tri(j, j)];
#pragma unroll
    for (int j = M - 1; j >= 0; --j) {
        double s = x[j];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k > j) s -= L[tri(k, j)] * x[k];
        x[j] = s;
    }
}

// y <- L^-1 y (forward substitution only)
template <int M>
UVS_DEV void chol_forward(const double (&L)[M * (M + 1) / 2], double (&x)[M]) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double s = x[j];
#pragma unroll
        for (int k = 0; k < M; ++k) if (k < j) s -= L[tri(j, k)] * x[k];
        x[j] = s * L[tri(j, j)];
    }
}

// Rank-revealing fallback of the IBVS pseudo-inverse: x = pinv(J) b with np.linalg.pinv's semantics
// (experiment.py:312: SVD, singular values <= 1e-15 sigma_max discarded) by one-sided (Hestenes) Jacobi ON J ITSELF, run
// cooperatively by the GS lanes of a group.  Working on J keeps the small singular values to high relative accuracy
// down to the numpy cutoff; an eigen-decomposition of the Gram matrix (what this function replaced) squares the
// condition number and cannot resolve sigma < ~1e-8 sigma_max, so it truncated where numpy still inverts.
//   W = J (MM x 6) for MM >= 6, W = J^T (6 x MM) for MM < 6:  R rows, C <= 6 columns.  Plane rotations from the right
//   orthogonalise the columns of A = W V (V accumulates them); at convergence A = U Sigma, so
//   MM >= 6:  x = V Sigma^-2 A^T b            MM < 6:  x = A Sigma^-2 V^T b          (columns with sigma > cutoff only).
// J: [MM x 6] row-major, b: MM doubles, both in shared memory; `work`: R*C + C*C doubles of per-group shared scratch
// (84 for MM = 8); x_out: 6 doubles (shared).  Every lane of the group must call; gl = lane index inside the group.
// J == 0 gives x = 0, as numpy does.  A rotation costs ~60 instructions per lane, a sweep C(C-1)/2 rotations, and the
// cyclic sweeps stop when one of them rotated nothing (typically 5-8 sweeps); rare by construction (cond(J) >~ 1e5).
template <int GS>
static __device__ __noinline__ void pinv_fallback_osj(int MM, const double* J, const double* b, double* work,
                                                      double* x_out, int gl) {
    const int lane = threadIdx.x & 31;
    const unsigned mask = (GS == 32) ? 0xffffffffu : (((1u << (GS & 31)) - 1u) << ((lane / GS) * GS));
    const bool tall = MM >= 6;
    const int R = tall ? MM : 6, C = tall ? 6 : MM;
    double* A = work;             // R x C, row-major (stride C)
    double* V = work + R * C;     // C x C
    __syncwarp(mask);
    for (int e = gl; e < R * C; e += GS) {
        const int r = e / C, c = e % C;
        A[e] = tall ? J[r * 6 + c] : J[c * 6 + r];
    }
    for (int e = gl; e < C * C; e += GS) V[e] = (e / C == e % C) ? 1.0 : 0.0;
    __syncwarp(mask);
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < C - 1; ++p)
            for (int q = p + 1; q < C; ++q) {
                double alpha = 0.0, beta = 0.0, gamma = 0.0;          // every lane: the same values (uniform control flow)
#pragma unroll 1
                for (int r = 0; r < R; ++r) {
                    const double ap = A[r * C + p], aq = A[r * C + q];
                    alpha += ap * ap; beta += aq * aq; gamma += ap * aq;
                }
                if (!(fabs(gamma) > 1.1102230246251565e-16 * sqrt(alpha * beta))) continue;   // also: NaN, zero columns
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(zeta * zeta + 1.0));
                const double c = rsqrt(t * t + 1.0), sn = t * c;
                __syncwarp(mask);                                     // all lanes have read columns p, q
#pragma unroll 1
                for (int i = gl; i < R + C; i += GS) {                // rows of A, then rows of V
                    double* row = (i < R) ? A + i * C : V + (i - R) * C;
                    const double vp = row[p], vq = row[q];
                    row[p] = c * vp - sn * vq;
                    row[q] = sn * vp + c * vq;
                }
                __syncwarp(mask);
            }
        if (!rotated) break;
    }
    // squared singular values, cutoff (numpy: s > rcond * s_max with rcond = 1e-15), projections.  Written with rolled
    // loops and NO local arrays: the register footprint of this noinline function is what its callers have to spill
    // around the call (measured: 70 registers here cost the block-diagonal closed loop 8-15 % of its throughput).
    double smax2 = 0.0;
    bool finite = true;                                   // np.linalg.pinv raises on NaN / Inf: hand NaN back (-> trial FAIL)
    for (int c = 0; c < C; ++c) {
        double a = 0.0;
        for (int r = 0; r < R; ++r) a += A[r * C + c] * A[r * C + c];
        finite = finite && (a < INFINITY);                // false for NaN too
        smax2 = fmax(smax2, a);
    }
    for (int r = 0; r < MM; ++r) finite = finite && (fabs(b[r]) < INFINITY);
    __syncwarp(mask);
    for (int i = gl; i < 6; i += GS) x_out[i] = 0.0;     // groups narrower than 6 lanes (GS = 2) own several components
    const double cut = 1e-15 * sqrt(smax2);
    for (int c = 0; c < C; ++c) {
        double s2 = 0.0, t = 0.0;
        for (int r = 0; r < R; ++r) s2 += A[r * C + c] * A[r * C + c];
        if (sqrt(s2) > cut && smax2 > 0.0) {                         // uniform in the group
            if (tall) {
                for (int r = 0; r < R; ++r) t += A[r * C + c] * b[r];
            } else {
                for (int k = 0; k < C; ++k) t += V[k * C + c] * b[k];
            }
            t /= s2;
            for (int i = gl; i < 6; i += GS) x_out[i] += (tall ? V[i * C + c] : A[i * C + c]) * t;   // each x_out[i] has one owner
        }
    }
    __syncwarp(mask);
    if (!finite) for (int i = gl; i < 6; i += GS) x_out[i] = NAN;
    __syncwarp(mask);
}

// doubles of per-group shared scratch pinv_fallback_osj needs for an M x 6 Jacobian with M >= 6 (M < 6 fits the group's
// own `sc` block)
__host__ __device__ constexpr int pinv_scratch_doubles(int M) { return (M >= 6) ? M * 6 + 36 : 0; }

// Cooperative minimum-norm least squares x = pinv(J) b  (experiment.py:312) by a group of GS lanes
// (GS = 32: one problem per warp; GS = 8 or 2: several problems per warp, the block-diagonal family).
// J: [M x 6] row-major in shared memory; b_lane: b_j on group lane j < M; sc: >= 56 doubles of
// per-group shared scratch; gl = lane index inside the group.  All 32 lanes of the warp must call.
// Fast path: L D L^T of the smaller Gram matrix (J^T J for M >= 6, J J^T for M < 6) + one step of
// iterative refinement on the residual (corrected semi-normal equations, error ~ eps cond(J)); if a
// pivot is <= 1e-11 trace(Gram) (cond(J) >~ 3e5; the refined normal equations are good to ~(eps cond^2)^2, i.e.
// 4e-10 there and 2e-6 at cond = 1e7, measured) the group runs pinv_fallback_osj (one-sided Jacobi on J, numpy's 1e-15
// sigma_max cutoff, error ~ eps cond).  Result in x on every lane of the group.  big: pinv_scratch_doubles(M) doubles of
// per-group shared scratch, required for M >= 6 (M < 6 uses the free head of `sc`).
template <int M, int GS>
UVS_DEV void pinv_group(const double* J, double b_lane, double* sc, int gl, double (&x)[6], int& flags,
                        double* big = nullptr) {
    double* sG = sc;          // 21
    double* sg = sc + 21;     // 6
    double* sb = sc + 27;     // 8
    double* sr = sc + 35;     // 8
    double* sx = sc + 43;     // 6
    constexpr int K = (M >= 6) ? 6 : M;
    __syncwarp();
    if (gl < M) sb[gl] = b_lane;
    __syncwarp();
    double L[K * (K + 1) / 2];
    double gmax = 0.0;
    if (M >= 6) {
        for (int e = gl; e < 27; e += GS) {
            if (e < 21) {
                // row of packed-lower entry e < 21, three bits each: 0 1 1 2 2 2 3 3 3 3 4 4 4 4 4 5 5 5 5 5 5
                const int i = (int)((0x5b6db2491b6d2448ull >> (3 * e)) & 7ull);
                const int j = e - i * (i + 1) / 2;
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < M; ++r) s += J[r * 6 + i] * J[r * 6 + j];
                sG[e] = s;
            } else {
                const int i = e - 21;
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < M; ++r) s += J[r * 6 + i] * sb[r];
                sg[i] = s;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < K * (K + 1) / 2; ++i) L[i] = sG[i];
    } else {
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
            for (int j = 0; j < K; ++j) if (j <= i) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < 6; ++c) s += J[i * 6 + c] * J[j * 6 + c];
                L[tri(i, j)] = s;
            }
    }
    // scale of the rank test: trace(G) (max diag <= trace <= K max diag; a sum is 1 DADD per term where fmax is ~6)
#pragma unroll
    for (int i = 0; i < K; ++i) gmax += L[tri(i, i)];
    bool ok = ldl_packed<K>(L);
#pragma unroll
    for (int j = 0; j < K; ++j) ok = ok && (L[tri(j, j)] * gmax * kPinvPivotTol < 1.0);        // L[j][j] = 1 / pivot
    // the fast path runs unconditionally (it only touches registers and this group's scratch) so
    // that the warp-wide barriers below are reached by every lane whatever the other groups do
    if (M >= 6) {
        double g[K];
#pragma unroll
        for (int i = 0; i < K; ++i) g[i] = sg[i];
        ldl_solve<K>(L, g);
        if (gl < M) {
            double s = b_lane;
#pragma unroll
            for (int c = 0; c < 6; ++c) s -= J[gl * 6 + c] * g[c];
            sr[gl] = s;
        }
        __syncwarp();
        if (gl < 6) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < M; ++r) s += J[r * 6 + gl] * sr[r];
            sx[gl] = s;
        }
        __syncwarp();
        double d[K];
#pragma unroll
        for (int i = 0; i < K; ++i) d[i] = sx[i];
        ldl_solve<K>(L, d);
        double dn2 = 0.0, xn2 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { x[i] = g[i] + d[i]; dn2 += d[i] * d[i]; xn2 += x[i] * x[i]; }
        ok = ok && (dn2 <= kPinvRefineTol2 * xn2);
    } else {
        double y[K];
#pragma unroll
        for (int i = 0; i < K; ++i) y[i] = sb[i];
        ldl_solve<K>(L, y);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < M; ++r) s += J[r * 6 + c] * y[r];
            x[c] = s;
        }
        double res[K];
#pragma unroll
        for (int r = 0; r < M; ++r) {
            double s = sb[r];
#pragma unroll
            for (int c = 0; c < 6; ++c) s -= J[r * 6 + c] * x[c];
            res[r] = s;
        }
        ldl_solve<K>(L, res);
        double dn2 = 0.0, xn2 = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < M; ++r) s += J[r * 6 + c] * res[r];
            x[c] += s;
            dn2 += s * s; xn2 += x[c] * x[c];
        }
        ok = ok && (dn2 <= kPinvRefineTol2 * xn2);
    }
    __syncwarp();
    if (!ok) {
        flags |= UVS_FLAG_RANK_DEFICIENT;
        int mm = M;
        asm volatile("" : "+r"(mm));       // opaque: keeps the compiler from cloning the noinline fallback per constant M
        pinv_fallback_osj<GS>(mm, J, sb, (M >= 6) ? big : sG, sx, gl);     // M < 6: sG (21 doubles) is unused, A + V need 6 M + M^2
    }
    __syncwarp();
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = sx[i];
    }
    __syncwarp();
}

template <int M>
UVS_DEV void pinv_warp(const double* J, double b_lane, double* sc, int lane, double (&x)[6], int& flags,
                       double* big = nullptr) {
    pinv_group<M, 32>(J, b_lane, sc, lane, x, flags, big);
}

}  // namespace uvs
