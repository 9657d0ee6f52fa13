// smallmat.cuh — register-resident small dense solves (fully unrolled, one thread).
// Used redundantly by every lane of a warp where a filter needs an m x m (m <= 8)
// factorisation: cheaper than a shuffle/smem-cooperative factorisation at this size
// because it needs no synchronisation and its ~8 dependent rsqrt steps overlap with
// other warps' DFMA streams.
#pragma once
#include "common.cuh"

namespace uvs {

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

// Slow, rank-revealing fallback of the IBVS pseudo-inverse (one thread, rolled loops, local arrays):
// x = pinv(J) b through the cyclic-Jacobi eigen-decomposition of the smaller Gram matrix, with
// eigenvalues <= 64 eps lambda_max discarded.  J [MM x 6] row-major.  J == 0 gives x = 0, as
// np.linalg.pinv does (experiment.py:312).
static __device__ __noinline__ void pinv_fallback(int MM, const double* J, const double* b, double* x) {
    const int K = MM >= 6 ? 6 : MM;
    double A[6][6], V[6][6], rhs[6], y[6];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            double s = 0.0;
            if (MM >= 6) for (int r = 0; r < MM; ++r) s += J[r * 6 + i] * J[r * 6 + j];
            else for (int c = 0; c < 6; ++c) s += J[i * 6 + c] * J[j * 6 + c];
            A[i][j] = s;
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int i = 0; i < K; ++i) {
        double s = 0.0;
        if (MM >= 6) for (int r = 0; r < MM; ++r) s += J[r * 6 + i] * b[r];
        else s = b[i];
        rhs[i] = s;
        y[i] = 0.0;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < K; ++i) {
            diag += A[i][i] * A[i][i];
            for (int j = 0; j < i; ++j) off += A[i][j] * A[i][j];
        }
        if (!(off > 1e-40 * diag)) break;
        for (int p = 0; p < K - 1; ++p)
            for (int q = p + 1; q < K; ++q) {
                const double apq = A[p][q];
                if (apq == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = rsqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < K; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < K; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < K; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    double lmax = 0.0;
    for (int i = 0; i < K; ++i) lmax = fmax(lmax, A[i][i]);
    const double cut = 64.0 * 2.220446049250313e-16 * lmax;
    for (int e = 0; e < K; ++e) {
        const double lam = A[e][e];
        if (!(lam > cut) || !(lmax > 0.0)) continue;
        double proj = 0.0;
        for (int i = 0; i < K; ++i) proj += V[i][e] * rhs[i];
        proj /= lam;
        for (int i = 0; i < K; ++i) y[i] += V[i][e] * proj;
    }
    if (MM >= 6) {
        for (int i = 0; i < 6; ++i) x[i] = y[i];
    } else {
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
            for (int r = 0; r < MM; ++r) s += J[r * 6 + c] * y[r];
            x[c] = s;
        }
    }
}

// The same fallback for the 6 x 6 Gram case (MM >= 6), run by the GS lanes of a group together with both work matrices in
// shared memory (`big`: 72 doubles, A then V) instead of one lane's local memory: the rotation sequence and every formula
// are those of pinv_fallback (same results), but the three 6-long update loops of a rotation are one step of six lanes,
// so it is ~4x faster — the 3-feature scene's square Jacobian passes near its singular configurations under impulsive
// noise with KF / IMCC-KF and this path then runs on a sizeable share of the frames.  rhs = J^T b is read from `g`
// (6 doubles, already computed by the fast path), the result goes to `x_out` (6 doubles, shared).  Every lane of the
// group must call; `gl` = lane index inside the group.
template <int GS>
static __device__ __noinline__ void pinv_fallback_coop(int MM, const double* J, const double* g, double* big, double* x_out,
                                                       int gl) {
    constexpr int K = 6;
    static_assert(GS >= K, "one lane per row of the Gram matrix");
    const int lane = threadIdx.x & 31;
    const unsigned mask = (GS == 32) ? 0xffffffffu : (((1u << (GS & 31)) - 1u) << ((lane / GS) * GS));
    double* A = big;
    double* V = big + 36;
    __syncwarp(mask);
    for (int e = gl; e < 36; e += GS) {
        const int i = e / 6, j = e % 6;
        double acc = 0.0;
        for (int r = 0; r < MM; ++r) acc += J[r * 6 + i] * J[r * 6 + j];
        A[e] = acc;
        V[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncwarp(mask);
    const int k = gl < K ? gl : K - 1;                 // lanes >= 6 shadow lane 5 and do not store
    const bool own = gl < K;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < K; ++i) {
            diag += A[i * 6 + i] * A[i * 6 + i];
            for (int j = 0; j < i; ++j) off += A[i * 6 + j] * A[i * 6 + j];
        }
        if (!(off > 1e-40 * diag)) break;              // uniform in the group: every lane read the same values
        for (int p = 0; p < K - 1; ++p)
            for (int q = p + 1; q < K; ++q) {
                const double apq = A[p * 6 + q];
                if (apq == 0.0) continue;              // uniform
                const double theta = (A[q * 6 + q] - A[p * 6 + p]) / (2.0 * apq);
                const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = rsqrt(t * t + 1.0), sn = t * c;
                // columns p, q of row k (A and V)
                const double akp = A[k * 6 + p], akq = A[k * 6 + q];
                const double vkp = V[k * 6 + p], vkq = V[k * 6 + q];
                __syncwarp(mask);                       // all reads of this rotation's inputs are done
                if (own) {
                    A[k * 6 + p] = c * akp - sn * akq;
                    A[k * 6 + q] = sn * akp + c * akq;
                    V[k * 6 + p] = c * vkp - sn * vkq;
                    V[k * 6 + q] = sn * vkp + c * vkq;
                }
                __syncwarp(mask);
                // rows p, q, column k
                const double apk = A[p * 6 + k], aqk = A[q * 6 + k];
                __syncwarp(mask);
                if (own) {
                    A[p * 6 + k] = c * apk - sn * aqk;
                    A[q * 6 + k] = sn * apk + c * aqk;
                }
                __syncwarp(mask);
            }
    }
    double lmax = 0.0;
    for (int i = 0; i < K; ++i) lmax = fmax(lmax, A[i * 6 + i]);
    const double cut = 64.0 * 2.220446049250313e-16 * lmax;
    double y = 0.0;
    for (int e = 0; e < K; ++e) {
        const double lam = A[e * 6 + e];
        if (!(lam > cut) || !(lmax > 0.0)) continue;
        double proj = 0.0;
        for (int i = 0; i < K; ++i) proj += V[i * 6 + e] * g[i];
        proj /= lam;
        y += V[k * 6 + e] * proj;
    }
    __syncwarp(mask);
    if (own) x_out[k] = y;
    __syncwarp(mask);
}

// Cooperative minimum-norm least squares x = pinv(J) b  (experiment.py:312) by a group of GS lanes
// (GS = 32: one problem per warp; GS = 8 or 2: several problems per warp, the block-diagonal family).
// J: [M x 6] row-major in shared memory; b_lane: b_j on group lane j < M; sc: >= 56 doubles of
// per-group shared scratch; gl = lane index inside the group.  All 32 lanes of the warp must call.
// Fast path: Cholesky of the smaller Gram matrix (J^T J for M >= 6, J J^T for M < 6) + one step of
// iterative refinement on the residual (corrected semi-normal equations, error ~ eps cond(J)); if a
// Cholesky pivot is <= 1e-13 trace(Gram) (cond(J) >~ 3e6) group lane 0 runs pinv_fallback.
// Result in x on every lane of the group.  big: optional 72 doubles of per-group shared scratch; when given (and
// M >= 6) the fallback runs cooperatively (pinv_fallback_coop) instead of on group lane 0.
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
    bool ok = chol_packed<K>(L);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double inv = L[tri(j, j)];
        ok = ok && (inv * inv * gmax * 1e-13 < 1.0);
    }
    // the fast path runs unconditionally (it only touches registers and this group's scratch) so
    // that the warp-wide barriers below are reached by every lane whatever the other groups do
    if (M >= 6) {
        double g[K];
#pragma unroll
        for (int i = 0; i < K; ++i) g[i] = sg[i];
        chol_solve<K>(L, g);
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
        chol_solve<K>(L, d);
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = g[i] + d[i];
    } else {
        double y[K];
#pragma unroll
        for (int i = 0; i < K; ++i) y[i] = sb[i];
        chol_solve<K>(L, y);
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
        chol_solve<K>(L, res);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < M; ++r) s += J[r * 6 + c] * res[r];
            x[c] += s;
        }
    }
    __syncwarp();
    if (!ok) {
        flags |= UVS_FLAG_RANK_DEFICIENT;
        if (M >= 6 && GS >= 6 && big != nullptr) pinv_fallback_coop<(GS >= 6 ? GS : 32)>(M, J, sg, big, sx, gl);
        else if (gl == 0) pinv_fallback(M, J, sb, sx);
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
