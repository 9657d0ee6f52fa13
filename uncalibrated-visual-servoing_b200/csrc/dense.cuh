// dense.cuh — one filter per warp, general SPD covariance.
//
// Replaces /root/reference/experiment.py:166-297 (predict, KF / MCKF / IMCC-KF / GMCKF=RMCKF
// update, Joseph covariance update) for B independent filters.
//
// Data layout (per warp; the filter never touches HBM between frames):
//   P    symmetric N x N, kept as its M(M+1)/2 lower 6x6 blocks.  The M(M-1)/2 OFF-DIAGONAL blocks
//        live in REGISTERS for the whole trial — lane l < NOFF owns block (bi, bj) = offdiag_of_lane(l),
//        36 doubles — and the M diagonal blocks live in shared memory (w.Pd, full 6x6 storage, both
//        triangles identical, kBlkStride = 36 doubles apart).
//   At   (P H^T)^T, Kt = K^T, Ut = U^T : M x N in shared memory, "k-major", so a lane's 6 consecutive
//        rows of one column are one 48-byte run (three LDS.128, broadcast across lanes sharing a
//        block row).
//   X    N state;  S, D  M x M (shared memory).
//
// Algebra (H = I_M (x) h^T, A = P^- H^T, S0 = H A):
//   KF      S = S0 + R          K = A S^-1            U = A
//   GMCKF   S = S0 + Rh         K = A S^-1            U = A + K (Rh - R)     Rh = Br (C + 1e-6 I)^-1 Br^T
//   IMCCKF  S = c S0 + R        K = c A S^-1          U = A + (1 - c) (A S^-1) R
//   then    P = P^- - K U^T   (== Joseph form (I-KH)P(I-KH)^T + K R K^T exactly, because A = K S
//           up to the factor c; rank-M update on the lower blocks only)
//   MCKF    fixed point on (Cx, Cy); the last gain K is not of the form A S^-1, so the Joseph form
//           is evaluated as the symmetric rank-2M update  P = P^- - K V^T - V K^T,  V = A - K (S0+R)/2.
//   joseph = 0 (tests/kalman_*): P = (I - K H) P^- = P^- - K A^T.
#pragma once
#include "common.cuh"
#include "smallmat.cuh"

// 1: the gain phase reads chol(S) from shared memory (frees 2 x M(M+1)/2 registers); 0: from registers
#ifndef UVS_DENSE_L_SMEM
#define UVS_DENSE_L_SMEM 1
#endif

namespace uvs {

// Distance (doubles) between the 6 x 6 diagonal blocks in shared memory.  36 = unpadded: a row of a block is 48 B = three 16-B
// bank groups, so eight consecutive rows (one per lane) fall on eight different bank groups whatever block they belong to
// (3 is coprime with 8); the padded 38 of round 1 made the rows of neighbouring blocks collide (measured: 36 is +2.0 % on the
// quad kernel, +0.9 % on the one-filter-per-warp kernel).
#ifndef UVS_BLK_STRIDE
#define UVS_BLK_STRIDE 36
#endif
constexpr int kBlkStride = UVS_BLK_STRIDE;

template <int M>
struct alignas(16) DenseConsts {          // per CTA
    double R[M * M];
    double Br[M * M];         // chol(R) lower, full storage
    double Brinv[M * M];      // inverse of Br (lower)
    int r_identity;
    int q_identity;
};

template <int M, bool MCKF>
struct alignas(16) DenseWS {
    static constexpr int N = 6 * M;
    static constexpr int LD = N + 2;           // row stride of At/Kt/Ut: 8(N+2) B is an odd number of 16-B bank groups
    static constexpr int NB = M * (M + 1) / 2;
    double Pd[M * kBlkStride];                  // diagonal blocks of P
    double At[M * LD];
    double Kt[M * LD];
    double Ut[M * LD];
    double X[N];
    double S[M * M];                            // S then D: >= 72 contiguous doubles for M >= 6, free once the gain is built:
    double D[M * M];                            // the work matrices of the cooperative pseudo-inverse fallback
    __device__ __forceinline__ double* jacobi_scratch() { return (M >= 6) ? S : nullptr; }
    static_assert(2 * M * M >= pinv_scratch_doubles(M), "S + D hold the work matrices of pinv_fallback_osj");
    double ib[56];                              // weights / IBVS scratch (G 21, g 6, b 8, res 8, x 6)
    double nu[M + (M & 1)];                     // innovation, read back by every lane
    double Lc[M * (M + 1) / 2 + ((M * (M + 1) / 2) & 1)];   // S = L D L^T: packed unit-lower factor, reciprocal pivots on the diagonal
    double q[6];                                // closed loop: joint positions, the last command,
    double dq[6];                               // and sin/cos of the DH angles of this frame
    double sc[12];
    double Lp[MCKF ? NB * kBlkStride : 2];      // chol(P), all lower blocks (MCKF only)
    double Xc[MCKF ? N : 2];
    double dx[MCKF ? N : 2];                    // X - Xc, then ex, then 1/cx
};

// this lane's off-diagonal block of P (lanes >= M(M-1)/2 carry an unused zero block)
struct PBlock {
    double v[6][6];
    int bi, bj;          // block row / column (bi > bj); lanes without a block carry (1, 0) and never use it
};

UVS_DEV int blk_index(int bi, int bj) { return bi * (bi + 1) / 2 + bj; }

// lane -> off-diagonal block (bi > bj)
UVS_DEV void offdiag_of_lane(int lane, int& bi, int& bj) {
    int i = 1;
    while ((i + 1) * i / 2 <= lane) ++i;        // smallest i with i(i+1)/2 > lane
    bi = i;
    bj = lane - i * (i - 1) / 2;
}

UVS_DEV void load_block(const double* p, double (&B)[6][6]) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        const double2 v = p2[i];
        B[(2 * i) / 6][(2 * i) % 6] = v.x;
        B[(2 * i + 1) / 6][(2 * i + 1) % 6] = v.y;
    }
}

UVS_DEV void store_block(double* p, const double (&B)[6][6]) {
    double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < 18; ++i)
        p2[i] = make_double2(B[(2 * i) / 6][(2 * i) % 6], B[(2 * i + 1) / 6][(2 * i + 1) % 6]);
}

UVS_DEV void load6(const double* p, double (&v)[6]) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < 3; ++i) { const double2 t = p2[i]; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
}

UVS_DEV void store6(double* p, const double (&v)[6]) {
    double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int i = 0; i < 3; ++i) p2[i] = make_double2(v[2 * i], v[2 * i + 1]);
}

// ---- P0 / P in and out --------------------------------------------------------------------------
// global P layouts: UVS_P_DENSE [N,N], UVS_P_BLOCKS [M,6,6]; src == nullptr -> identity.
// The lower triangle of a dense P is what is read (the diagonal blocks are symmetrised from it).
template <int M>
UVS_DEV void load_P(PBlock& pb, double* Pd, const double* src, int layout, int lane) {
    constexpr int N = 6 * M, LD = N + 2, NOFF = M * (M - 1) / 2;
    int bi = 1, bj = 0;
    if (lane < NOFF) offdiag_of_lane(lane, bi, bj);
    pb.bi = bi; pb.bj = bj;
    const bool dense = (src != nullptr) && (layout == UVS_P_DENSE) && (lane < NOFF);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) pb.v[r][c] = dense ? src[(size_t)(6 * bi + r) * N + 6 * bj + c] : 0.0;
    for (int e = lane; e < M * 36; e += 32) {
        const int d = e / 36, rc = e % 36, r = rc / 6, c = rc % 6;
        const int hi = r >= c ? r : c, lo = r >= c ? c : r;
        double v;
        if (src == nullptr) v = (r == c) ? 1.0 : 0.0;
        else if (layout == UVS_P_DENSE) v = src[(size_t)(6 * d + hi) * N + 6 * d + lo];
        else v = src[d * 36 + hi * 6 + lo];
        Pd[d * kBlkStride + rc] = v;
    }
}

template <int M>
UVS_DEV void store_P(const PBlock& pb, const double* Pd, double* dst, int layout, int lane) {
    constexpr int N = 6 * M, LD = N + 2, NOFF = M * (M - 1) / 2;
    if (layout == UVS_P_DENSE) {
        if (lane < NOFF) {
            const int bi = pb.bi, bj = pb.bj;
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    dst[(size_t)(6 * bi + r) * N + 6 * bj + c] = pb.v[r][c];
                    dst[(size_t)(6 * bj + c) * N + 6 * bi + r] = pb.v[r][c];
                }
        }
        for (int e = lane; e < M * 36; e += 32) {
            const int d = e / 36, rc = e % 36;
            dst[(size_t)(6 * d + rc / 6) * N + 6 * d + rc % 6] = Pd[d * kBlkStride + rc];
        }
    } else {
        for (int e = lane; e < M * 36; e += 32) dst[e] = Pd[(e / 36) * kBlkStride + e % 36];
    }
}

// ---- phase 1: P <- P + Q ;  At <- (P H^T)^T ------------------------------------------------------
template <int M>
UVS_DEV void predict_and_A(PBlock& pb, double* Pd, double* At, const double (&h)[6], const double* Qg,
                           int q_identity, int lane) {
    constexpr int N = 6 * M, LD = N + 2, NOFF = M * (M - 1) / 2;
    if (lane < NOFF) {
        const int bi = pb.bi, bj = pb.bj;
        if (!q_identity) {
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) pb.v[r][c] += __ldg(Qg + (size_t)(6 * bi + r) * N + 6 * bj + c);
        }
        // rows 6bi.. of column bj, and (by symmetry) rows 6bj.. of column bi
        double lo[6], hi[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double a = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) a += pb.v[r][c] * h[c];
            lo[r] = a;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) a += pb.v[r][c] * h[r];
            hi[c] = a;
        }
        store6(At + bj * LD + 6 * bi, lo);
        store6(At + bi * LD + 6 * bj, hi);
    }
    for (int row = lane; row < N; row += 32) {
        const int d = row / 6, r = row % 6;
        double* pr = Pd + d * kBlkStride + r * 6;
        double v[6];
        if (q_identity) {
            pr[r] += 1.0;            // one dynamic-address RMW; the row load below is ordered after it (same thread)
            load6(pr, v);
        } else {
            load6(pr, v);
#pragma unroll
            for (int c = 0; c < 6; ++c) { v[c] += __ldg(Qg + (size_t)row * N + 6 * d + c); pr[c] = v[c]; }
        }
        double a = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) a += v[c] * h[c];
        At[d * LD + row] = a;
    }
}

// ---- phase 5: P <- P - L1 R1^T (- L2 R2^T) on the lower blocks ------------------------------------
// RANK2 = false: P -= Kt^T-rows x Ut^T-rows  (P[i][j] -= sum_k K[i][k] U[j][k])
// RANK2 = true : P[i][j] -= sum_k K[i][k] V[j][k] + V[i][k] K[j][k]      (V held in Ut)
// Off-diagonal part for k in [K0, K1).  Branch-free: lanes without a block (>= NOFF) carry the dummy block
// (1, 0) and update registers nobody reads, so the FMAs can be scheduled into a neighbouring dependency chain
// (see ibvs_with_rank_update) instead of sitting behind a divergent branch.
template <int M, bool RANK2, int K0, int K1>
UVS_DEV void rank_update_offdiag(PBlock& pb, const double* Kt, const double* Ut) {
    constexpr int N = 6 * M, LD = N + 2;
    const int bi = pb.bi, bj = pb.bj;
#pragma unroll
    for (int k = K0; k < K1; ++k) {
        double ki[6], uj[6];
        load6(Kt + k * LD + 6 * bi, ki);
        load6(Ut + k * LD + 6 * bj, uj);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) pb.v[r][c] -= ki[r] * uj[c];
        if (RANK2) {
            double vi[6], kj[6];
            load6(Ut + k * LD + 6 * bi, vi);
            load6(Kt + k * LD + 6 * bj, kj);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) pb.v[r][c] -= vi[r] * kj[c];
        }
    }
}

// Diagonal blocks: 4 lanes per block.  Lane s owns row rowA = 5 - s (columns 0..rowA) and, for s = 1, 2,
// row rowB = s - 1 (columns 0..rowB): 6 + 6 + 6 + 3 = 21 lower entries.  All four lanes load the block
// row's six U values (broadcast) and their own two K values, and run 6 + 2 FMAs per k on fixed
// registers; entries outside the lane's set are computed and dropped.
template <int M, bool RANK2>
UVS_DEV void rank_update_diag(double* Pd, const double* Kt, const double* Ut, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    const int d = lane >> 2, s = lane & 3;
    const bool on = (M == 8) || (d < M);       // M = 8: all 32 lanes, no branch
    const int rowA = 5 - s;
    const int rowB = (s == 2) ? 1 : 0;
    double* pd = Pd + (on ? d : 0) * kBlkStride;
    double accA[6], accB[2];
    if (on) {
        const double* kt = Kt + 6 * d;
        const double* ut = Ut + 6 * d;
        load6(pd + rowA * 6, accA);
        accB[0] = pd[rowB * 6]; accB[1] = pd[rowB * 6 + 1];
#pragma unroll
        for (int k = 0; k < M; ++k) {
            double u[6];
            load6(ut + k * LD, u);
            const double kA = kt[k * LD + rowA], kB = kt[k * LD + rowB];
#pragma unroll
            for (int e = 0; e < 6; ++e) accA[e] -= kA * u[e];
            accB[0] -= kB * u[0];
            accB[1] -= kB * u[1];
            if (RANK2) {
                double kk[6];
                load6(kt + k * LD, kk);
                const double uA = ut[k * LD + rowA], uB = ut[k * LD + rowB];
#pragma unroll
                for (int e = 0; e < 6; ++e) accA[e] -= uA * kk[e];
                accB[0] -= uB * kk[0];
                accB[1] -= uB * kk[1];
            }
        }
    }
    // every lane has read its rows (including the mirrored upper-triangle entries it computes and drops) before any
    // lane writes them: without this barrier compute-sanitizer's racecheck reports those unused reads against the
    // mirror writes of the neighbouring lanes
    __syncwarp();
    if (on) {
#pragma unroll
        for (int e = 0; e < 6; ++e)
            if (e <= rowA) { pd[rowA * 6 + e] = accA[e]; pd[e * 6 + rowA] = accA[e]; }
        if (s == 1 || s == 2) {
            pd[rowB * 6] = accB[0]; pd[rowB] = accB[0];
            if (s == 2) { pd[rowB * 6 + 1] = accB[1]; pd[6 + rowB] = accB[1]; }
        }
    }
}

template <int M, bool RANK2>
UVS_DEV void rank_update(PBlock& pb, double* Pd, const double* Kt, const double* Ut, int lane) {
    rank_update_offdiag<M, RANK2, 0, M>(pb, Kt, Ut);
    rank_update_diag<M, RANK2>(Pd, Kt, Ut, lane);
}

// IBVS law for M >= 6 (x = pinv(J) b, smallmat.cuh pinv_group with one problem per warp) with the pending
// covariance update P -= K U^T executed INSIDE it.  The pseudo-inverse is a chain of short dependent stages
// (Gram entries -> 6x6 Cholesky -> solve -> residual -> solve) that leaves the FP64 pipe mostly idle; the
// rank-M update is 36 independent FMAs per lane and per k that depend only on K, U (shared memory) and the
// lane's P block.  Each stage below is one basic block holding both, so the scheduler fills the chain's
// latency gaps with update FMAs.  Same results as rank_update() followed by pinv_warp().
template <int M>
UVS_DEV void ibvs_with_rank_update(PBlock& pb, double* Pd, const double* Kt, const double* Ut, const double* J,
                                   double b_lane, double* sc, int lane, double (&x)[6], int& flags, double* big UVS_CLK_ARGS) {
    static_assert(M >= 6, "M < 6 uses the J J^T path of pinv_group");
    constexpr int K = 6;
    constexpr int C1 = M / 4, C2 = (3 * M) / 4, C3 = M - 1;      // k ranges handed to the four stages
    double* sG = sc;          // 21
    double* sg = sc + 21;     // 6
    double* sb = sc + 27;     // 8
    double* sr = sc + 35;     // 8
    double* sx = sc + 43;     // 6
    __syncwarp();
    if (lane < M) sb[lane] = b_lane;
    __syncwarp();
    // The stages are written branch-free (clamped indices, predicated stores) so that each one stays a single
    // basic block together with its share of the rank update.
    {   // stage 1: Gram matrix J^T J (21 entries) and J^T b (6), one per lane: sc[e] = sum_r J[r][i] * q[r]
        const int e = lane < 27 ? lane : 26;
        const bool gram = e < 21;
        const int i = gram ? (int)((0x5b6db2491b6d2448ull >> (3 * e)) & 7ull) : e - 21;
        const int j = e - i * (i + 1) / 2;
        const double* q = gram ? J + j : sb;         // second factor: column j of J (stride 6) or b (stride 1)
        const int qs = gram ? 6 : 1;
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < M; ++r) s += J[r * 6 + i] * q[r * qs];
        rank_update_offdiag<M, false, 0, C1>(pb, Kt, Ut);
        if (lane < 27) sc[e] = s;                    // sG = sc[0..21), sg = sc[21..27)
    }
    __syncwarp();
    UVS_CLK(6);
    double L[K * (K + 1) / 2], g[K];
    bool ok;
    {   // stage 2: factor, first solve, residual
#pragma unroll
        for (int i = 0; i < K * (K + 1) / 2; ++i) L[i] = sG[i];
#pragma unroll
        for (int i = 0; i < K; ++i) g[i] = sg[i];
        double gsum = 0.0;
#pragma unroll
        for (int i = 0; i < K; ++i) gsum += L[tri(i, i)];
        ok = ldl_packed<K>(L);
#pragma unroll
        for (int j = 0; j < K; ++j) ok = ok && (L[tri(j, j)] * gsum * kPinvPivotTol < 1.0);    // L[j][j] = 1 / pivot
        ldl_solve<K>(L, g);
        const int row = lane < M ? lane : M - 1;
        double s = sb[row];
#pragma unroll
        for (int c = 0; c < 6; ++c) s -= J[row * 6 + c] * g[c];
        rank_update_offdiag<M, false, C1, C2>(pb, Kt, Ut);
        if (lane < M) sr[lane] = s;
    }
    __syncwarp();
    UVS_CLK(7);
    {   // stage 3: J^T r
        const int col = lane < 6 ? lane : 5;
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < M; ++r) s += J[r * 6 + col] * sr[r];
        rank_update_offdiag<M, false, C2, C3>(pb, Kt, Ut);
        if (lane < 6) sx[lane] = s;
    }
    __syncwarp();
    UVS_CLK(8);
    {   // stage 4: refinement solve
        double d[K];
#pragma unroll
        for (int i = 0; i < K; ++i) d[i] = sx[i];
        ldl_solve<K>(L, d);
        double dn2 = 0.0, xn2 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) { x[i] = g[i] + d[i]; dn2 += d[i] * d[i]; xn2 += x[i] * x[i]; }
        ok = ok && (dn2 <= kPinvRefineTol2 * xn2);        // a posteriori conditioning test, see smallmat.cuh
        rank_update_offdiag<M, false, C3, M>(pb, Kt, Ut);
        rank_update_diag<M, false>(Pd, Kt, Ut, lane);
    }
    __syncwarp();
    UVS_CLK(9);
    if (!ok) {
        flags |= UVS_FLAG_RANK_DEFICIENT;
        int mm = M;
        asm volatile("" : "+r"(mm));       // see pinv_group
        pinv_fallback_osj<32>(mm, J, sb, big, sx, lane);
    }
    __syncwarp();
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = sx[i];
    }
    __syncwarp();
}

// ---- MCKF helpers: block Cholesky of P and the products with it ---------------------------------
// Lp <- chol(P) in block layout; diagonal blocks hold the 6x6 factor (lower, TRUE diagonal, upper
// part zeroed).  Returns false if P is not numerically SPD.
template <int M>
UVS_DEV bool block_cholesky(const PBlock& pb, const double* Pd, double* Lp, int lane) {
    constexpr int NOFF = M * (M - 1) / 2;
    if (lane < NOFF) {
        store_block(Lp + blk_index(pb.bi, pb.bj) * kBlkStride, pb.v);
    }
    for (int e = lane; e < M * 36; e += 32) Lp[blk_index(e / 36, e / 36) * kBlkStride + e % 36] = Pd[(e / 36) * kBlkStride + e % 36];
    __syncwarp();
    bool ok = true;
    for (int bj = 0; bj < M; ++bj) {
        // (1) factor the diagonal block redundantly in registers
        double Ld[21];
        {
            const double* pd = Lp + blk_index(bj, bj) * kBlkStride;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j <= i) Ld[tri(i, j)] = pd[i * 6 + j];
        }
        ok = chol_packed<6>(Ld) && ok;          // reciprocal diagonal
        __syncwarp();
        for (int e = lane; e < 36; e += 32) {
            double* pd = Lp + blk_index(bj, bj) * kBlkStride;
            const int i = e / 6, j = e % 6;
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = 0; b < 6; ++b)
                    if (b <= a && a == i && b == j) v = (a == b) ? 1.0 / Ld[tri(a, b)] : Ld[tri(a, b)];
            pd[e] = v;
        }
        // (2) panel: rows of the blocks below, x <- x L_jj^-T  (forward substitution per row)
        for (int rr = lane; rr < (M - 1 - bj) * 6; rr += 32) {
            const int bi = bj + 1 + rr / 6, r = rr % 6;
            double* pr = Lp + blk_index(bi, bj) * kBlkStride + r * 6;
            double x[6];
            load6(pr, x);
            chol_forward<6>(Ld, x);
#pragma unroll
            for (int c = 0; c < 6; ++c) pr[c] = x[c];
        }
        __syncwarp();
        // (3) trailing update: B_ik -= B_ij B_kj^T for bj < bk <= bi
        const int nt = (M - 1 - bj) * (M - bj) / 2;
        if (lane < nt) {
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= lane) ++ti;
            const int tk = lane - ti * (ti + 1) / 2;
            const int bi = bj + 1 + ti, bk = bj + 1 + tk;
            double C[6][6];
            const double* pi = Lp + blk_index(bi, bj) * kBlkStride;
            const double* pk = Lp + blk_index(bk, bj) * kBlkStride;
            double* pc = Lp + blk_index(bi, bk) * kBlkStride;
            load_block(pc, C);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                double ci[6], ck[6];
#pragma unroll
                for (int r = 0; r < 6; ++r) { ci[r] = pi[r * 6 + k]; ck[r] = pk[r * 6 + k]; }
#pragma unroll
                for (int r = 0; r < 6; ++r)
#pragma unroll
                    for (int c = 0; c < 6; ++c) C[r][c] -= ci[r] * ck[c];
            }
            store_block(pc, C);
        }
        __syncwarp();
    }
    return ok;
}

// y <- Bp^-1 y  (block forward substitution, y in shared memory, length N)
template <int M>
UVS_DEV void block_forward(const double* Lp, double* y, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    for (int bj = 0; bj < M; ++bj) {
        const double* pd = Lp + blk_index(bj, bj) * kBlkStride;
        double yj[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double s = y[6 * bj + i];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < i) s -= pd[i * 6 + k] * yj[k];
            yj[i] = s / pd[i * 6 + i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (lane == i) y[6 * bj + i] = yj[i];
        for (int row = 6 * (bj + 1) + lane; row < N; row += 32) {
            const double* pr = Lp + blk_index(row / 6, bj) * kBlkStride + (row % 6) * 6;
            double s = y[row];
#pragma unroll
            for (int k = 0; k < 6; ++k) s -= pr[k] * yj[k];
            y[row] = s;
        }
        __syncwarp();
    }
}

// Ahat^T (into Kt) <- (Bp diag(icx) Bp^T H^T)^T ; T1^T staged in Ut.  icx[j] = 1/cx_j.
template <int M>
UVS_DEV void mckf_weighted_A(const double* Lp, const double* icx, const double (&h)[6], double* T1t,
                             double* Ahat_t, int lane) {
    constexpr int N = 6 * M, LD = N + 2, NB = M * (M + 1) / 2;
    for (int e = lane; e < M * LD; e += 32) T1t[e] = 0.0;
    __syncwarp();
    // T1[6bj+c][k] = icx * sum_r Lp[6k+r][6bj+c] h_r   over blocks (k, bj), bj <= k
    for (int b = lane; b < NB; b += 32) {
        int k = 0;
        while ((k + 1) * (k + 2) / 2 <= b) ++k;
        const int bj = b - k * (k + 1) / 2;
        double B[6][6];
        load_block(Lp + b * kBlkStride, B);        // diagonal blocks have their upper part zeroed
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) s += B[r][c] * h[r];
            T1t[k * LD + 6 * bj + c] = s * icx[6 * bj + c];
        }
    }
    __syncwarp();
    // Ahat[6bi+r][k] = sum_{bj <= min(bi,k)} sum_c Lp[6bi+r][6bj+c] T1[6bj+c][k]
    for (int e = lane; e < M * M; e += 32) {
        const int bi = e / M, k = e % M;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        const int lim = bi < k ? bi : k;
        for (int bj = 0; bj <= lim; ++bj) {
            double B[6][6], t[6];
            load_block(Lp + blk_index(bi, bj) * kBlkStride, B);
            load6(T1t + k * LD + 6 * bj, t);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[r] += B[r][c] * t[c];
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) Ahat_t[k * LD + 6 * bi + r] = acc[r];
    }
    __syncwarp();
}

// ---- S = s0_scale * H Ain + Rh, factorised redundantly in every lane's registers ------------------
// Ain: A^T-like array (M x N).  w_lane: lane i < M carries the i-th diagonal weight of W, where
// Rh = Br W Br^T (plain_R: Rh = R).  Optionally stores the matrix D used by U = A + Kp D:
// d_mode 1 -> D = Rh - R,  d_mode 2 -> D = dscale * R.  With R = I both are diagonal and only
// ddiag[k] = D[k][k] (shared, M doubles) is written.
template <int M>
UVS_DEV bool build_and_factor_S(const double* Ain, const double (&h)[6], const DenseConsts<M>& cc,
                                double w_lane, bool plain_R, double s0_scale, double* Ssm, double* wsm,
                                double* Dsm, double* ddiag, int d_mode, double dscale,
                                double (&L)[M * (M + 1) / 2], int lane, bool r_id) {
    constexpr int N = 6 * M, LD = N + 2;
    if (lane < M) {
        wsm[lane] = w_lane;
        if (d_mode != 0 && r_id) ddiag[lane] = (d_mode == 1) ? w_lane - 1.0 : dscale;
    }
    __syncwarp();
    for (int e = lane; e < M * M; e += 32) {
        const int i = e / M, k = e % M;
        double a[6];
        load6(Ain + k * LD + 6 * i, a);
        double s0 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s0 += h[r] * a[r];
        double rh;
        if (plain_R) rh = r_id ? ((i == k) ? 1.0 : 0.0) : cc.R[e];
        else if (r_id) rh = (i == k) ? wsm[i] : 0.0;
        else {
            rh = 0.0;
#pragma unroll
            for (int j = 0; j < M; ++j) rh += cc.Br[i * M + j] * wsm[j] * cc.Br[k * M + j];
        }
        Ssm[e] = s0_scale * s0 + rh;
        if (!r_id) {
            if (d_mode == 1) Dsm[e] = rh - cc.R[e];
            else if (d_mode == 2) Dsm[e] = dscale * cc.R[e];
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j <= i) L[tri(i, j)] = Ssm[i * M + j];
    return ldl_packed<M>(L);          // unit lower factor, reciprocal pivots on the diagonal
}

// x_r <- (L D L^T)^-1 x_r for R right-hand sides at once, with the factor (ldl_packed layout: unit lower, reciprocal
// pivots on the diagonal) read from shared memory: every access is a warp-wide broadcast with a compile-time offset, so
// the factor costs no registers in the gain phase, and the R substitutions are independent dependency chains the
// scheduler can interleave.  Unit diagonal: one FMA per substitution step on the critical path.
template <int M, int R>
UVS_DEV void ldl_solve_sm(const double* Ls, double (&x)[R][M]) {
#pragma unroll
    for (int j = 0; j < M; ++j) {
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k < j) {
                const double l = Ls[tri(j, k)];
#pragma unroll
                for (int r = 0; r < R; ++r) x[r][j] -= l * x[r][k];
            }
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double d = Ls[tri(j, j)];
#pragma unroll
        for (int r = 0; r < R; ++r) x[r][j] *= d;
    }
#pragma unroll
    for (int j = M - 1; j >= 0; --j) {
#pragma unroll
        for (int k = 0; k < M; ++k)
            if (k > j) {
                const double l = Ls[tri(k, j)];
#pragma unroll
                for (int r = 0; r < R; ++r) x[r][j] -= l * x[r][k];
            }
    }
}

struct StepOut {
    int flags = 0;
    int epoch = 0;
    bool rank_pending = false;
};

// One predict + update of the filter held in (pb, w).  z_lane: Z_i on lane i < M.  Returns kappa_i on
// lane i < M (experiment.py:306-308).  K^T is left in w.Kt.
// DEFER = true (M >= 6, not MCKF): the covariance update is left pending (out.rank_pending) for the caller to
// run inside ibvs_with_rank_update(); K^T and U^T stay valid in w.Kt / w.Ut until then.
// FAST = true: Q = I, R = I and the Joseph form are compile-time facts (the reference's hard-coded
// experiment.py:73-76,297), so the general-Q/R code and its warp-uniform branches are not in the frame loop.
template <int M, int METHOD, bool DEFER = false, bool FAST = false>
UVS_DEV double dense_step(PBlock& pb, DenseWS<M, METHOD == UVS_METHOD_MCKF>& w, const DenseConsts<M>& cc,
                          const uvs_filter_params& fp, const double (&h)[6], double z_lane, double sigma,
#ifdef UVS_PHASE_CLOCKS
                          int lane, StepOut& out, double inv_s2 UVS_CLK_ARGS) {
#else
                          int lane, StepOut& out, double inv_s2 = 0.0) {
#endif
    constexpr int N = 6 * M, LD = N + 2;
    const bool r_id = FAST || (cc.r_identity != 0);
    const bool joseph = FAST || (fp.joseph != 0);
    predict_and_A<M>(pb, w.Pd, w.At, h, fp.Q, FAST ? 1 : cc.q_identity, lane);
    UVS_CLK(2);

    // innovation nu = Z - H X (lane i < M), published through shared memory
    double nu_l = 0.0;
    if (lane < M) {
        double x[6];
        load6(w.X + 6 * lane, x);
        nu_l = z_lane;
#pragma unroll
        for (int r = 0; r < 6; ++r) nu_l -= h[r] * x[r];
        w.nu[lane] = nu_l;
    }
    __syncwarp();
    // whitened innovation e = Br^-1 nu
    double e_l = 0.0;
    if (lane < M) {
        if (r_id) e_l = nu_l;
        else {
#pragma unroll
            for (int j = 0; j < M; ++j) e_l += cc.Brinv[lane * M + j] * w.nu[j];
        }
    }
    double kappa = 1.0;
    double L[M * (M + 1) / 2];

    if (METHOD != UVS_METHOD_MCKF) {
        double gain_scale = 1.0, dscale = 0.0;
        bool plain_R = true;
        double s0_scale = 1.0, w_lane = 1.0;
        int d_mode = 0;
        if (METHOD == UVS_METHOD_GMCKF) {
            // experiment.py:274-276, 280.  FAST (closed loop): the caller supplies 1 / sigma^2 from its per-CTA table, the
            // weight is exp(-(e^2 / 2) (1 / sigma^2)) (argument within 1 ulp of utils.py:171-172's) and 1 / (c + 1e-6)
            // is the branch-free reciprocal (c + 1e-6 lies in [1e-6, 1.000001])
            const double c = (lane < M) ? (FAST ? exp(-0.5 * (e_l * e_l) * inv_s2) : gaussian_kernel(e_l, sigma)) : 1.0;
            kappa = c;
            w_lane = FAST ? fast_div(1.0, c + 1e-6) : 1.0 / (c + 1e-6);
            plain_R = false;
            d_mode = joseph ? 1 : 0;
        } else if (METHOD == UVS_METHOD_IMCCKF) {
            double q = 0.0;                                                        // experiment.py:258-261
#pragma unroll
            for (int j = 0; j < M; ++j) { const double ej = shfl_d(e_l, j); q += ej * ej; }
            const double c = gaussian_kernel(sqrt(q), sigma);
            s0_scale = c; gain_scale = c; dscale = 1.0 - c;
            d_mode = joseph ? 2 : 0;
        }
        double* ddiag = w.ib + 8;
        UVS_CLK(3);
        const bool ok = build_and_factor_S<M>(w.At, h, cc, w_lane, plain_R, s0_scale, w.S, w.ib, w.D, ddiag, d_mode,
                                              dscale, L, lane, r_id);
        UVS_CLK(4);
        if (!ok) {                       // S not SPD: skip the correction, P keeps +Q
            out.flags |= UVS_FLAG_NOT_SPD;
            for (int e = lane; e < M * LD; e += 32) w.Kt[e] = 0.0;
            __syncwarp();
            return kappa;
        }
        const bool use_D = d_mode != 0;
        const bool diag_D = r_id;
#if UVS_DENSE_L_SMEM
        {   // every lane holds the same factor: publish it once, then let the registers go
            constexpr int NL = M * (M + 1) / 2;
            if (lane == 0) {                                  // one writer: a same-address STS.128 from 32 lanes costs 4 wavefronts
#pragma unroll
                for (int i = 0; i + 1 < NL; i += 2) *reinterpret_cast<double2*>(w.Lc + i) = make_double2(L[i], L[i + 1]);
                if (NL & 1) w.Lc[NL - 1] = L[NL - 1];
            }
            __syncwarp();
        }
#endif
        double nu[M];
#pragma unroll
        for (int j = 0; j < M; j += 2) {
            const double2 t = *reinterpret_cast<const double2*>(w.nu + j);
            nu[j] = t.x; nu[j + 1] = t.y;
        }
        // Gain rows.  N > 32: lanes 0..N/2-1 take rows (lane, lane + N/2) and run the two triangular solves as
        // one interleaved instruction stream (two independent dependency chains, each factor entry read once).
        constexpr int RPL = (N > 32) ? 2 : 1, NROWL = N / RPL;
        if (lane < NROWL) {
            double a[RPL][M], x[RPL][M];
#pragma unroll
            for (int j = 0; j < RPL; ++j)
#pragma unroll
                for (int k = 0; k < M; ++k) { a[j][k] = w.At[k * LD + lane + j * NROWL]; x[j][k] = a[j][k]; }
#if UVS_DENSE_L_SMEM
            {
                const double* Ls = w.Lc;
                asm volatile("" : "+l"(Ls));                      // keep the loads inside the block (no hoisting into registers)
                ldl_solve_sm<M, RPL>(Ls, x);
            }
#else
#pragma unroll
            for (int j = 0; j < RPL; ++j) ldl_solve<M>(L, x[j]);   // Kp row = A row S^-1
#endif
#pragma unroll
            for (int j = 0; j < RPL; ++j) {
                const int row = lane + j * NROWL;
                double dxr = 0.0;
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    w.Kt[k * LD + row] = gain_scale * x[j][k];
                    dxr += x[j][k] * nu[k];
                }
                w.X[row] += gain_scale * dxr;                    // experiment.py:193 / 265 / 291
                if (use_D && diag_D) {
#pragma unroll
                    for (int k = 0; k < M; ++k) w.Ut[k * LD + row] = a[j][k] + x[j][k] * ddiag[k];
                } else {
#pragma unroll
                    for (int k = 0; k < M; ++k) {
                        double u = a[j][k];
                        if (use_D) {
#pragma unroll
                            for (int jj = 0; jj < M; ++jj) u += x[j][jj] * w.D[jj * M + k];
                        }
                        w.Ut[k * LD + row] = u;
                    }
                }
            }
        }
        __syncwarp();
        UVS_CLK(5);
        if (DEFER) { out.rank_pending = true; return kappa; }
        rank_update<M, false>(pb, w.Pd, w.Kt, w.Ut, lane);
        __syncwarp();
        return kappa;
    } else {
        // ------------------------------------------------------------------ MCKF (experiment.py:194-250)
        // The first pass of the fixed point starts from X_corrected = X, so its state residuals are zero, Cx = I and
        // P_hat = P: it needs neither chol(P) nor the weighted product, and its gain K = A (S0 + R_hat)^-1 has the
        // form that lets the Joseph update collapse to the rank-M form P -= K U^T, U = A + K (R_hat - R).  With the
        // reference's fpi_threshold = 0.1 ~96 % of the frames stop after that pass, so the 48 x 48 block
        // Cholesky (N^3 / 3 flops, the dominant cost of this branch) is factored lazily, by the first frame that
        // needs a second pass; the reference factors P every frame (experiment.py:203) but only uses the factor there.
        double nu[M];
#pragma unroll
        for (int j = 0; j < M; j += 2) {
            const double2 t = *reinterpret_cast<const double2*>(w.nu + j);
            nu[j] = t.x; nu[j + 1] = t.y;
        }
        for (int row = lane; row < N; row += 32) w.Xc[row] = w.X[row];
        __syncwarp();
        double* ddiag = w.ib + 8;
        bool skip = false;
        int epoch = 0;
        double difference = INFINITY;
        while (difference > fp.fpi_threshold && epoch < fp.fpi_epoch_max) {
            const double* Ain = w.At;
            if (epoch > 0) {
                if (epoch == 1) {                                       // Bp = chol(P^-), experiment.py:203
                    const bool spd = block_cholesky<M>(pb, w.Pd, w.Lp, lane);
                    if (!spd) { out.flags |= UVS_FLAG_NOT_SPD; skip = true; break; }
                }
                // ex = Bp^-1 (X - Xc) ; cx = G(ex)
                for (int row = lane; row < N; row += 32) w.dx[row] = w.X[row] - w.Xc[row];
                __syncwarp();
                block_forward<M>(w.Lp, w.dx, lane);
                bool cx_zero = false;
                for (int row = lane; row < N; row += 32) {
                    const double cx = gaussian_kernel(w.dx[row], sigma);
                    cx_zero = cx_zero || (cx == 0.0);
                    w.dx[row] = 1.0 / cx;
                }
                cx_zero = __any_sync(UVS_FULL_MASK, cx_zero);
                __syncwarp();
                if (cx_zero) { out.flags |= UVS_FLAG_CX_SINGULAR; skip = true; break; }
                mckf_weighted_A<M>(w.Lp, w.dx, h, w.Ut, w.Kt, lane);
                Ain = w.Kt;
            }
            // ey = Br^-1 (Z - H Xc) ; cy = G(ey)
            double r_l = 0.0;
            if (lane < M) {
                r_l = z_lane;
#pragma unroll
                for (int r = 0; r < 6; ++r) r_l -= h[r] * w.Xc[6 * lane + r];
            }
            double ey = 0.0;
            if (r_id) ey = r_l;
            else {
#pragma unroll
                for (int j = 0; j < M; ++j) { const double rj = shfl_d(r_l, j); if (lane < M) ey += cc.Brinv[lane * M + j] * rj; }
            }
            const double cy = (lane < M) ? gaussian_kernel(ey, sigma) : 1.0;
            if (__any_sync(UVS_FULL_MASK, cy == 0.0)) {                 // experiment.py:231-236
                out.flags |= UVS_FLAG_SKIP_SINGULAR; skip = true; break;
            }
            // first pass: also D = R_hat - R for the rank-M Joseph form (w.D / ddiag)
            const bool ok = build_and_factor_S<M>(Ain, h, cc, 1.0 / cy, false, 1.0, w.S, w.ib, w.D, ddiag,
                                                  epoch == 0 ? 1 : 0, 0.0, L, lane, r_id);
            if (!ok) { out.flags |= UVS_FLAG_NOT_SPD; skip = true; break; }
            // K rows, X_corrected = X + K (Z - H X)        experiment.py:238-242
            double num = 0.0, den = 0.0;
            for (int row = lane; row < N; row += 32) {
                double x[M];
#pragma unroll
                for (int k = 0; k < M; ++k) x[k] = Ain[k * LD + row];
                ldl_solve<M>(L, x);
                double xc = w.X[row];
#pragma unroll
                for (int k = 0; k < M; ++k) { w.Kt[k * LD + row] = x[k]; xc += x[k] * nu[k]; }
                const double old = w.Xc[row];
                num += (xc - old) * (xc - old);
                den += old * old;
                w.Xc[row] = xc;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { num += shfl_xor_d(num, o); den += shfl_xor_d(den, o); }
            difference = sqrt(num) / sqrt(den);                        // experiment.py:244
            ++epoch;
            __syncwarp();
            if (epoch == fp.fpi_epoch_max && fp.mckf_skip) {           // experiment.py:246-248
                out.flags |= UVS_FLAG_SKIP_MAXEPOCH; skip = true;
            }
        }
        out.epoch = epoch;
        if (skip) return kappa;
        if (epoch == 1) {
            // X = X_corrected ; one pass: K = A S^-1, so P = P^- - K U^T with U = A + K (R_hat - R)
            for (int row = lane; row < N; row += 32) {
                w.X[row] = w.Xc[row];
                double kr[M];
#pragma unroll
                for (int k = 0; k < M; ++k) kr[k] = w.Kt[k * LD + row];
#pragma unroll
                for (int k = 0; k < M; ++k) {
                    double u = w.At[k * LD + row];
                    if (r_id) u += kr[k] * ddiag[k];
                    else {
#pragma unroll
                        for (int j = 0; j < M; ++j) u += kr[j] * w.D[j * M + k];
                    }
                    w.Ut[k * LD + row] = u;
                }
            }
            __syncwarp();
            rank_update<M, false>(pb, w.Pd, w.Kt, w.Ut, lane);
            __syncwarp();
            return kappa;
        }
        // several passes: the last gain is not of the form A S^-1; Joseph as the symmetric rank-2M update with
        // V = A - K (S0 + R)/2, S0 = H A of the TRUE predicted covariance
        for (int e = lane; e < M * M; e += 32) {
            const int i = e / M, k = e % M;
            double s0 = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) s0 += h[r] * w.At[k * LD + 6 * i + r];
            w.D[e] = s0;
        }
        __syncwarp();
        for (int row = lane; row < N; row += 32) {
            w.X[row] = w.Xc[row];
            double kr[M];
#pragma unroll
            for (int k = 0; k < M; ++k) kr[k] = w.Kt[k * LD + row];
#pragma unroll
            for (int k = 0; k < M; ++k) {
                double v = w.At[k * LD + row];
#pragma unroll
                for (int j = 0; j < M; ++j) v -= 0.5 * kr[j] * (w.D[j * M + k] + (r_id ? (j == k ? 1.0 : 0.0) : cc.R[j * M + k]));
                w.Ut[k * LD + row] = v;
            }
        }
        __syncwarp();
        rank_update<M, true>(pb, w.Pd, w.Kt, w.Ut, lane);
        __syncwarp();
        return kappa;
    }
}

// Fill the per-CTA constants (R, chol(R), its inverse) — experiment.py:76,121-122.  One thread.
template <int M>
__device__ void init_consts(DenseConsts<M>& cc, const uvs_filter_params& fp) {
    bool ident = (fp.R == nullptr);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) cc.R[i * M + j] = fp.R ? fp.R[i * M + j] : (i == j ? 1.0 : 0.0);
    if (!ident) {
        ident = true;
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) ident = ident && (cc.R[i * M + j] == (i == j ? 1.0 : 0.0));
    }
    cc.r_identity = ident;
    cc.q_identity = (fp.Q == nullptr);
    // Cholesky (lower) and inverse of the factor
    for (int i = 0; i < M * M; ++i) { cc.Br[i] = 0.0; cc.Brinv[i] = 0.0; }
    for (int j = 0; j < M; ++j) {
        double d = cc.R[j * M + j];
        for (int k = 0; k < j; ++k) d -= cc.Br[j * M + k] * cc.Br[j * M + k];
        const double s = sqrt(d);
        cc.Br[j * M + j] = s;
        for (int i = j + 1; i < M; ++i) {
            double v = cc.R[i * M + j];
            for (int k = 0; k < j; ++k) v -= cc.Br[i * M + k] * cc.Br[j * M + k];
            cc.Br[i * M + j] = v / s;
        }
    }
    for (int c = 0; c < M; ++c)
        for (int i = c; i < M; ++i) {
            double v = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k) v -= cc.Br[i * M + k] * cc.Brinv[k * M + c];
            cc.Brinv[i * M + c] = v / cc.Br[i * M + i];
        }
}

// Per-frame trace of a closed-loop kernel (uvs_closed_loop_io, ABI 4): the state / covariance / command / weights of
// frame k of trial b.  Called after the covariance update of the frame has completed (it is deferred into the IBVS
// pseudo-inverse on the fast path).  `traced` is warp-uniform and false in production launches.
UVS_DEV bool io_traced(const uvs_closed_loop_io& io) {
    return io.X_log != nullptr || io.P_log != nullptr || io.dq_log != nullptr || io.kappa_log != nullptr || io.flags_log != nullptr;
}

template <int M>
UVS_DEV void trace_frame(const uvs_closed_loop_io& io, size_t bt, const double* Pd, const double* X,
                         double dq_l, double kappa, int flags, int lane) {
    constexpr int N = 6 * M;
    if (io.X_log) for (int r = lane; r < N; r += 32) io.X_log[bt * N + r] = X[r];
    if (io.P_log)           // diagonal blocks only (UVS_P_BLOCKS): a copy out of shared memory, no register traffic
        for (int e = lane; e < M * 36; e += 32) io.P_log[bt * (M * 36) + e] = Pd[(e / 36) * kBlkStride + e % 36];
    if (io.dq_log && lane < 6) io.dq_log[bt * 6 + lane] = dq_l;
    if (io.kappa_log && lane < M) io.kappa_log[bt * M + lane] = kappa;
    if (io.flags_log && lane == 0) io.flags_log[bt] = flags;
}

}  // namespace uvs
