// dense_quad.cuh — the warp-specialised closed loop of the dense family (general SPD covariance): FOUR filters per CTA.
//
// Why: a frame of the dense closed loop is one long dependent chain (plant -> noise -> innovation -> weights -> S ->
// factorisation -> gain -> state -> pseudo-inverse -> command -> plant ...; ~9 800 cycles for a warp that has a scheduler
// to itself, measured with UVS_DENSE_CTAS_PER_SM=1), and the one-filter-per-warp kernel (dense_kernels.cuh) can keep only
// 8 of those chains per SM: the filter's covariance (72 registers per lane) and the register arrays of the chain phases
// (the 8x8 factor: 72, the 6x6 factor + solves: ~80) live in the SAME warp, so it needs 255 registers and every register
// less is a spill on the chain (10 warps / SM at 200 registers: 0.58x, measured).
//
// Here the two register populations live in different warps of a 5-warp CTA:
//   warps 0..3  "covariance warps": warp g holds the off-diagonal blocks of filter g's P in registers for the whole trial
//               and does the WIDE work of the frame: P += Q, A = P H^T, S0 = H A (from the registers that hold P_ik h),
//               X += A y, the Gram matrix J^T J and J^T b of the IBVS normal equations (one entry per lane), P -= K U^T;
//               and, through one non-inlined call every 32 / M frames, the Philox / alpha-stable draws of the next 32 / M
//               frames (inlined, the samplers made ptxas park P in local memory for the whole frame loop);
//   warp 4      "chain warp": group g = lanes 8g..8g+7 runs the SCALAR chains of filter g, one measurement row per lane,
//               the four filters of the CTA in lockstep: plant (sincos, point in camera, pixel), hold logic, innovation,
//               kernel weights, S = S0 + R_hat and its L D L^T, the explicit S^-1 (one column per lane; the reference forms
//               it too: K = P H^T inv(S), experiment.py:192,286-287), y = g S^-1 nu, the gain rows K = g A S^-1 and
//               U = A + (A S^-1) D, the pseudo-inverse from the Gram matrix, command, finiteness, error integrals, log,
//               trace — one instruction stream for four filters instead of four.
// Neither role needs more than 128 registers, so THREE CTAs = 12 filters are resident per SM (the shared-memory work
// space of a filter is unchanged, 14.5 KB), and the chain instructions are issued once per four filters.
//
// The roles exchange through the filter's shared-memory work space (DenseWS, as before) and meet at four named barriers
// per frame; a role that only publishes at a barrier ARRIVES and goes on, the role that consumes WAITS:
//   (4) dq, alive, done        | covariance: wait;  predict, A, S0        | chain: arrive;  log, integrals, integrate, NEXT
//                              |                                          |        frame's plant, hold, innovation, weights
//   (1) A, S0 / nu, weights, b | covariance: arrive;  (noise batch)       | chain: wait;  S, L D L^T, S^-1, y
//   (2) S^-1, y, ok            | covariance: wait;  X += A y, Gram, J^T b | chain: arrive;  gain rows K, U
//   (3) X, Gram / K, U         | covariance: wait;  P -= K U^T            | chain: wait;  pinv, command, finiteness
// Barrier (4) sits right after the command.  A trial that ends because a feature left the image is found by the chain warp
// after the covariance warp has started that frame's predict; the covariance warp restores the diagonal of P it saved (w.D).
// Arithmetic: dense_step<M, METHOD, DEFER, FAST> + ibvs_with_rank_update (dense.cuh) with two re-associations inside the
// 1e-9 gate: the gain rows come from the explicit S^-1 (not two substitutions per row), and X += A (g S^-1 nu) (not
// (g A S^-1) nu).  tests/test_gpu_parity.py: the replay gate against experiment.py's own run, and
// test_quad_kernel_matches_one_filter_per_warp_kernel.  DESIGN.md section 3.1 has the measurements behind every choice.
//
// Scope: the FAST configuration (Q = R = NULL, Joseph form, no scene plant: experiment.py:73-76,297 as hard-coded by the
// reference), M >= 6, KF / IMCC-KF / GMCKF.  Everything else stays on dense_closed_loop_kernel.
#pragma once
#include "dense_kernels.cuh"
#include <cstdio>

// register cap: 3 CTAs x 5 warps x 32 lanes x 128 registers (136 leaves room for 2 CTAs only: allocation granularity, measured)
#ifndef UVS_QUAD_MAXNREG
#define UVS_QUAD_MAXNREG 128
#endif
#define UVS_QUAD_BOUNDS __maxnreg__(UVS_QUAD_MAXNREG)
#ifndef UVS_QUAD_GAIN_UNROLL
#define UVS_QUAD_GAIN_UNROLL 2
#endif
#ifndef UVS_QUAD_GAIN_RB
#define UVS_QUAD_GAIN_RB 2
#endif
#ifndef UVS_QUAD_RANK_UNROLL
#define UVS_QUAD_RANK_UNROLL 2
#endif

// profiling build (-DUVS_PHASE_CLOCKS, tools/phase_clocks.py --quad): wall cycles per role and phase, slots 12.. of the buffer
#ifdef UVS_PHASE_CLOCKS
#define QCLK_DECL(n) unsigned long long qacc[n] = {}; long long qlast = clock64();
#define QCLK(i) do { const long long now_ = clock64(); qacc[i] += (unsigned long long)(now_ - qlast); qlast = now_; } while (0)
#define QCLK_FLUSH(n, base) do { if ((threadIdx.x & 31) == 0) for (int i_ = 0; i_ < (n); ++i_) atomicAdd(a.phase_clk + (base) + i_, qacc[i_]); } while (0)
#else
#define QCLK_DECL(n)
#define QCLK(i) do { } while (0)
#define QCLK_FLUSH(n, base) do { } while (0)
#endif

namespace uvs {

constexpr int kQuadFilters = 4;
constexpr int kQuadGainUnroll = UVS_QUAD_GAIN_UNROLL, kQuadRankUnroll = UVS_QUAD_RANK_UNROLL;
constexpr int kQuadThreads = (kQuadFilters + 1) * 32;

struct alignas(16) QuadCtl {           // per filter
    double noise[2][32];               // fresh draws of 32 / M frames, double-buffered by batch: [batch & 1][slot * M + coordinate]
    int ok;                            // S factorised (else: NOT_SPD, the correction is skipped)
    int alive;                         // the trial is still running
    int done;                          // the frame before barrier (4) completed (trace rows are written for it)
};

// the CTA's dynamic shared memory as seen by either role (declared inside each function so that every access is an LDS /
// STS on the shared window, not a generic load through a pointer argument)
#define QUAD_SMEM_VIEW(MM)                                                                                        \
    extern __shared__ __align__(16) unsigned char smem_raw[];                                                      \
    unsigned char* ws_base = smem_raw + sigma_table_bytes(a.cfg.T);                                                \
    QuadCtl* ctlv = reinterpret_cast<QuadCtl*>(ws_base + kQuadFilters * quad_ws_stride<MM>());                     \
    volatile int* quad_alive = reinterpret_cast<volatile int*>(ctlv + kQuadFilters);
#define QUAD_WS(MM, g) (*reinterpret_cast<DenseWS<MM, false>*>(ws_base + (g) * quad_ws_stride<MM>()))

// Named barriers of the CTA.  quad_bar(): everybody waits (trial start / end).  The four per-frame barriers (ids 1..4)
// are split: a role that only PUBLISHES at a barrier arrives without waiting (bar.arrive) and goes on; the role that
// CONSUMES waits (bar.sync).  Covariance warps: arrive (1), wait (2) (3) (4); chain warp: wait (1) (3), arrive (2) (4).
// The warp reconverges (__syncwarp) before it touches a CTA barrier, and the barrier instructions are the non-.aligned
// forms: after lane-dependent loops the lanes of a warp are not guaranteed to arrive together (compute-sanitizer synccheck).
UVS_DEV void quad_bar() {
    __syncwarp();
    asm volatile("barrier.sync 5, %0;" ::"n"(kQuadThreads) : "memory");
}
template <int ID> UVS_DEV void quad_wait() {
    __syncwarp();
    asm volatile("barrier.sync %0, %1;" ::"n"(ID), "n"(kQuadThreads) : "memory");
}
template <int ID> UVS_DEV void quad_arrive() {
    __syncwarp();
    __threadfence_block();
    asm volatile("barrier.arrive %0, %1;" ::"n"(ID), "n"(kQuadThreads) : "memory");
}

UVS_DEV bool quad_group_any(bool p, int lane) {
    const unsigned b = __ballot_sync(UVS_FULL_MASK, p);
    return ((b >> (lane & ~7)) & 0xffu) != 0u;
}

UVS_DEV double quad_group_sum(double v) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// Distance between the work spaces of the CTA's four filters: == 64 (mod 128) bytes.  The chain warp's group g reads
// filter g's arrays with the same offsets in all four groups (8 lanes x 8 B each): with this stride groups 0 / 1 and 2 / 3
// fall on disjoint halves of the 32 banks and a 64-bit access is its ideal 2 wavefronts (sizeof(DenseWS<8>) == 16 mod
// 128 put all four groups on overlapping banks: 237 wavefronts per filter-step in the gain phase where 147 are needed, ncu).
template <int M>
__host__ __device__ constexpr size_t quad_ws_stride() {
    return (sizeof(DenseWS<M, false>) + 127) / 128 * 128 + 64;
}

template <int M>
__host__ __device__ constexpr size_t quad_smem_bytes(int T) {
    return sigma_table_bytes(T) + kQuadFilters * (quad_ws_stride<M>() + sizeof(QuadCtl)) + 16;
}

// Gain rows by the CHAIN warp: lane gl of group g computes rows gl, gl + 8, ..., of filter g's K = gain_scale A S^-1,
// RB rows at a time (independent accumulator sets: the chain warp holds no P, it has the registers), and
// U = A + (A S^-1) diag(ddiag).  The covariance warps do not take part: with P in their registers ptxas serialises this
// phase on four temporaries (LDS -> DFMA -> LDS on the same registers, measured 6 900 cycles per frame for ~1 000 of work).
template <int M, int RB>
UVS_DEV void quad_gain_rows_chain(DenseWS<M, false>& w, double gain_scale, bool use_D, int gl) {
    constexpr int N = 6 * M, LD = N + 2;
    const double* __restrict__ At = w.At;
    const double* __restrict__ Si = w.S;
    const double* __restrict__ ddiag = w.Lc + 8;      // w.Lc: y [0..M), ddiag [8..8+M) (w.ib is the Gram matrix's, written meanwhile)
    double* __restrict__ Kt = w.Kt;
    double* __restrict__ Ut = w.Ut;
#pragma unroll 1
    for (int r0 = gl; r0 < N; r0 += 8 * RB) {
        int rowv[RB];                                            // rows of this batch (clamped; rows >= N are not stored)
#pragma unroll
        for (int t = 0; t < RB; ++t) rowv[t] = (r0 + 8 * t < N) ? r0 + 8 * t : r0;
        double kx[RB][M], av[RB][M];                             // rows of A are kept for U (no second read)
#pragma unroll
        for (int t = 0; t < RB; ++t)
#pragma unroll
            for (int k = 0; k < M; ++k) { kx[t][k] = 0.0; av[t][k] = At[k * LD + rowv[t]]; }
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const double2* sr = reinterpret_cast<const double2*>(Si + j * M);      // row j of S^-1
#pragma unroll
            for (int k = 0; k < M; k += 2) {
                const double2 s2 = sr[k / 2];
#pragma unroll
                for (int t = 0; t < RB; ++t) { kx[t][k] += av[t][j] * s2.x; kx[t][k + 1] += av[t][j] * s2.y; }
            }
        }
        // epilogue: every load before the first store (the compiler orders shared loads behind earlier shared stores it
        // cannot prove disjoint: written row by row this was LDS -> DFMA -> STS -> LDS ..., each pair a full shared-memory
        // round trip, and took half of the chain warp's frame — ncu source view)
        {
            double dd[M];
#pragma unroll
            for (int k = 0; k < M; ++k) dd[k] = ddiag[k];
            if (use_D) {
#pragma unroll
                for (int t = 0; t < RB; ++t)
#pragma unroll
                    for (int k = 0; k < M; ++k) av[t][k] = av[t][k] + kx[t][k] * dd[k];        // row of U
            }
        }
#pragma unroll
        for (int t = 0; t < RB; ++t) {
            const int row = r0 + 8 * t;
            if (row < N) {
#pragma unroll
                for (int k = 0; k < M; ++k) Kt[k * LD + row] = gain_scale * kx[t][k];
                if (use_D) {
#pragma unroll
                    for (int k = 0; k < M; ++k) Ut[k * LD + row] = av[t][k];
                }
            }
        }
    }
}

// P <- P + Q (Q = I), A^T = (P H^T)^T and the lower triangle of S0 = H A by the covariance warp: predict_and_A() of dense.cuh
// plus build_and_factor_S()'s S0, with S0[bi][bj] = h . (P_bibj h) taken from the registers that hold P_bibj h (the same sum,
// in the same order, as reading A back from shared memory: 24 wavefronts per filter-step less).  The diagonal of P before
// the increment is kept in w.D: restored if the chain warp finds that the trial ended (feature out of view) BEFORE this
// frame — it evaluates the plant while this warp predicts.
template <int M>
UVS_DEV void quad_predict(PBlock& pb, DenseWS<M, false>& w, const double (&h)[6], int lane) {
    constexpr int N = 6 * M, LD = N + 2, NOFF = M * (M - 1) / 2;
    double* __restrict__ At = w.At;
    double* __restrict__ S0 = w.S;
    double* __restrict__ Pd = w.Pd;
    double* __restrict__ Dsave = w.D;
    if (lane < NOFF) {
        const int bi = pb.bi, bj = pb.bj;
        double lo[6], hi[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 6; ++c) acc += pb.v[r][c] * h[c];
            lo[r] = acc;
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r) acc += pb.v[r][c] * h[r];
            hi[c] = acc;
        }
        store6(At + bj * LD + 6 * bi, lo);
        store6(At + bi * LD + 6 * bj, hi);
        double s0 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s0 += h[r] * lo[r];
        S0[bi * M + bj] = s0;
    }
    for (int row = lane; row < N; row += 32) {
        const int d = row / 6, r = row % 6;
        double* pr = Pd + d * kBlkStride + r * 6;
        Dsave[row] = pr[r];
        pr[r] += 1.0;
        double v[6];
        load6(pr, v);
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += v[c] * h[c];
        At[d * LD + row] = acc;
    }
    __syncwarp();
    if (lane < M) {
        double av[6];
        load6(At + lane * LD + 6 * lane, av);
        double s0 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) s0 += h[r] * av[r];
        S0[lane * M + lane] = s0;
    }
}

// X += A y by the covariance warp (y in w.Lc[0..M), left there by the chain warp)
template <int M>
UVS_DEV void quad_state_update(DenseWS<M, false>& w, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    const double* __restrict__ At = w.At;
    const double* __restrict__ yv = w.Lc;
    double* __restrict__ X = w.X;
    double y[M];
#pragma unroll
    for (int j = 0; j < M; j += 2) {
        const double2 t = *reinterpret_cast<const double2*>(yv + j);
        y[j] = t.x; y[j + 1] = t.y;
    }
    for (int row = lane; row < N; row += 32) {
        double a[M];
#pragma unroll
        for (int j = 0; j < M; ++j) a[j] = At[j * LD + row];
        const double x0 = X[row];
        double dx = 0.0;
#pragma unroll
        for (int j = 0; j < M; ++j) dx += a[j] * y[j];
        X[row] = x0 + dx;
    }
}

// Gram matrix J^T J (21 packed entries) and J^T b (6) of the IBVS normal equations, one entry per lane of the covariance
// warp (stage 1 of ibvs_with_rank_update, dense.cuh): sc[e] = sum_r J[r][i] q[r].  b was left in sc[27..27+M) by the chain
// warp before barrier (1).
template <int M>
UVS_DEV void quad_gram(const double* __restrict__ J, double* __restrict__ sc, int lane) {
    const double* sb = sc + 27;
    const int e = lane < 27 ? lane : 26;
    const bool gram = e < 21;
    const int i = gram ? (int)((0x5b6db2491b6d2448ull >> (3 * e)) & 7ull) : e - 21;
    const int j = e - i * (i + 1) / 2;
    const double* q = gram ? J + j : sb;         // second factor: column j of J (stride 6) or b (stride 1)
    const int qs = gram ? 6 : 1;
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < M; ++r) s += J[r * 6 + i] * q[r * qs];
    if (lane < 27) sc[e] = s;                    // sG = sc[0..21), sg = sc[21..27)
}

// The rest of the pseudo-inverse by the chain warp's 8-lane groups: pinv_group<M, 8> (smallmat.cuh) from the point where
// the Gram matrix and J^T b are in sc.  x = pinv(J) b on every lane of the group.
template <int M>
UVS_DEV void quad_pinv_from_gram(const double* J, double b_lane, double* sc, int gl, double (&x)[6], int& flags, double* big) {
    constexpr int K = 6;
    double* sG = sc;          // 21
    double* sg = sc + 21;     // 6
    double* sb = sc + 27;     // 8
    double* sr = sc + 35;     // 8
    double* sx = sc + 43;     // 6
    double L[K * (K + 1) / 2], g[K];
#pragma unroll
    for (int i = 0; i < K * (K + 1) / 2; ++i) L[i] = sG[i];
#pragma unroll
    for (int i = 0; i < K; ++i) g[i] = sg[i];
    double gmax = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i) gmax += L[tri(i, i)];
    bool ok = ldl_packed<K>(L);
#pragma unroll
    for (int j = 0; j < K; ++j) ok = ok && (L[tri(j, j)] * gmax * kPinvPivotTol < 1.0);        // L[j][j] = 1 / pivot
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
    __syncwarp();
    if (!ok) {
        flags |= UVS_FLAG_RANK_DEFICIENT;
        int mm = M;
        asm volatile("" : "+r"(mm));       // see pinv_group
        pinv_fallback_osj<8>(mm, J, sb, big, sx, gl);
    }
    __syncwarp();
    if (!ok) {
#pragma unroll
        for (int i = 0; i < 6; ++i) x[i] = sx[i];
    }
    __syncwarp();
}

// Fresh draws of frames k0 .. k0 + 32/M - 1 of one filter, one per lane of its covariance warp (lane = slot * M +
// coordinate).  NOT inlined: the samplers (log / tan / pow slow paths) want ~100 registers; as a call the warp saves what
// it must of P around ONE call site that runs every 32 / M frames, in the shadow of the chain warp's pseudo-inverse.
template <int M>
static __device__ __noinline__ void quad_noise_batch(const uvs_noise_params& np, double rho, double alpha, uint64_t trial,
                                                     int k0, int T, double* out, int lane) {
    constexpr int FB = 32 / M;
    const int fs = lane / M;
    if (fs < FB && k0 + fs < T) out[lane] = noise_sample(np, rho, alpha, trial, (uint32_t)(k0 + fs), (uint32_t)(lane % M));
}

// P -= K U^T on the lower blocks by the covariance warp: rank_update() of dense.cuh with the k loop of the off-diagonal
// part ROLLED.  This warp is not on the frame's critical path (the chain warp is), so it trades the instruction-level
// parallelism of the unrolled form for registers: unrolled, ptxas hoists the loads of several k above the FMAs and parks
// P itself in local memory for the whole frame loop (measured in the SASS: 72 LDL.64 + 36 STL.64 per frame).
template <int M>
UVS_DEV void quad_rank_update(PBlock& pb, double* Pd, const double* Kt, const double* Ut, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    const double* __restrict__ kp = Kt + 6 * pb.bi;
    const double* __restrict__ up = Ut + 6 * pb.bj;
#pragma unroll kQuadRankUnroll
    for (int k = 0; k < M; ++k) {
        double ki[6], uj[6];
        load6(kp + k * LD, ki);
        load6(up + k * LD, uj);
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) pb.v[r][c] -= ki[r] * uj[c];
    }
    rank_update_diag<M, false>(Pd, Kt, Ut, lane);
}

// ---- role 1: the covariance warp of filter `warp` --------------------------------------------------------------
// The two roles are separate NON-INLINED functions on purpose: ptxas allocates registers per function, and with both
// roles inlined into one kernel body it parked P in local memory for the whole frame loop although a third of the
// register file was unused in that branch (seen in the SASS).
template <int M, int METHOD>
static __device__ __noinline__ void quad_covariance_role(const ClosedLoopArgs& a, int warp) {
    QUAD_SMEM_VIEW(M)
    constexpr int N = 6 * M, LD = N + 2;
    using WS = DenseWS<M, false>;
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const int T = cfg.T;
    const int lane = threadIdx.x & 31;
    const bool philox = (a.io.noise_in == nullptr) && (cfg.noise.type != UVS_NOISE_NONE);
    const bool traced = io_traced(a.io);
    constexpr bool use_D = METHOD != UVS_METHOD_KF;          // Joseph with R_hat != R (GMCKF) / c != 1 (IMCC-KF)
    constexpr int FB = 32 / M;                               // frames of noise drawn at once by a covariance warp
    // =================================================================== covariance warp of filter `warp`
    WS& w = QUAD_WS(M, warp);
    QuadCtl& ctl = ctlv[warp];
    volatile int* v_alive = &ctl.alive;
    volatile int* v_ok = &ctl.ok;
    volatile int* v_done = &ctl.done;
    for (long long b0 = (long long)blockIdx.x * kQuadFilters; b0 < a.B; b0 += (long long)gridDim.x * kQuadFilters) {
        const long long b = b0 + warp;
        const bool trial_ok = b < a.B;
        const long long bb = trial_ok ? b : 0;
        const uint64_t trial = a.trial0 + (uint64_t)bb * a.stride;
        double rho = cfg.noise.rho, alpha = cfg.noise.alpha;             // noise level of this trial (main.py:121-127)
        if (cfg.sweep_len > 0 && cfg.sweep_param != 0) {
            long long lvl = (long long)(trial / (uint64_t)(cfg.sweep_epoch > 0 ? cfg.sweep_epoch : 1));
            if (lvl >= cfg.sweep_len) lvl = cfg.sweep_len - 1;
            if (cfg.sweep_param == 1) rho = cfg.sweep[lvl]; else alpha = cfg.sweep[lvl];
        }
        if (a.io.X0_in)
            for (int r = lane; r < N; r += 32) w.X[r] = a.io.X0_in[bb * N + r];
        if (philox) quad_noise_batch<M>(cfg.noise, rho, alpha, trial, 0, T, ctl.noise[0], lane);
        PBlock pb;
        load_P<M>(pb, w.Pd, cfg.P0, UVS_P_DENSE, lane);
        quad_bar();                                                      // init: X, dq, alive, noise published
        QCLK_DECL(6)
        for (int k = 0; k < T; ++k) {
            // the block coordinates and the lane are made opaque once per frame: otherwise every shared-memory address derived
            // from them is hoisted out of the frame loop as a loop invariant, and ptxas keeps ~40 such addresses in
            // registers while it parks half of P in local memory (seen in the SASS)
            int ln = lane;
            asm volatile("" : "+r"(pb.bi), "+r"(pb.bj), "+r"(ln));
            bool live = *v_alive != 0;
            if (live) {
                double h[6];                                             // experiment.py:183-188: the previous command
                load6(w.dq, h);
                quad_predict<M>(pb, w, h, ln);
            }
            QCLK(0);
            quad_arrive<1>();                                            // (1) A, S0 published
            // the draws of the next FB frames, into the other half of the noise buffer, while the chain warp finishes the
            // measurement and factorises S
            if (philox && (k + 1) % FB == 0 && k + 1 < T)
                quad_noise_batch<M>(cfg.noise, rho, alpha, trial, k + 1, T, ctl.noise[((k + 1) / FB) & 1], ln);
            QCLK(1);
            quad_wait<2>();                                              // (2) S^-1, y published
            QCLK(2);
            if (live && *v_alive == 0) {                                 // ended out of view: undo the predict
                for (int row = ln; row < N; row += 32) w.Pd[(row / 6) * kBlkStride + (row % 6) * 7] = w.D[row];
                live = false;
            }
            const bool ok = live && (*v_ok != 0);
            // while the chain warp forms the gain rows K = A S^-1 and U: the state correction X += A y, y = gain_scale S^-1 nu
            // (experiment.py:193 / 265 / 291: X + K nu = X + A (S^-1 nu)), then, with J = X, the Gram matrix and J^T b
            if (ok) quad_state_update<M>(w, ln);
            __syncwarp();
            quad_gram<M>(w.X, w.ib, ln);
            quad_wait<3>();                                              // (3) X, Gram | K, U
            QCLK(3);
            if (ok) {
                quad_rank_update<M>(pb, w.Pd, w.Kt, use_D ? w.Ut : w.At, ln);
                __syncwarp();
            }
            QCLK(4);
            quad_wait<4>();                                              // (4) dq, alive published
            QCLK(5);
            if (traced && live && (*v_done != 0) && trial_ok) {          // state / covariance after the update of frame k
                const size_t bt = (size_t)b * T + k;
                if (a.io.X_log) for (int r = ln; r < N; r += 32) a.io.X_log[bt * N + r] = w.X[r];
                if (a.io.P_log)
                    for (int e = ln; e < M * 36; e += 32) a.io.P_log[bt * (M * 36) + e] = w.Pd[(e / 36) * kBlkStride + e % 36];
            }
            if (*quad_alive == 0) break;
        }
        QCLK_FLUSH(6, 22);
        if (trial_ok) {
            if (a.io.X_out) for (int r = lane; r < N; r += 32) a.io.X_out[b * N + r] = w.X[r];
            if (a.io.P_out) {
                const size_t psz = (a.io.p_layout == UVS_P_DENSE) ? (size_t)N * N : (size_t)M * 36;
                store_P<M>(pb, w.Pd, a.io.P_out + b * psz, a.io.p_layout, lane);
            }
        }
        quad_bar();                                                      // work space free for the next four trials
    }
}

// ---- role 2: the chain warp: group g = filter g, lane gl = measurement row ---------------------------------------
template <int M, int METHOD>
static __device__ __noinline__ void quad_chain_role(const ClosedLoopArgs& a) {
    QUAD_SMEM_VIEW(M)
    const double2* sig_tab = (sigma_table_bytes(a.cfg.T) != 0) ? reinterpret_cast<const double2*>(smem_raw) : nullptr;
    constexpr int N = 6 * M, LD = N + 2;
    using WS = DenseWS<M, false>;
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const int T = cfg.T;
    const int lane = threadIdx.x & 31;
    const bool philox = (a.io.noise_in == nullptr) && (cfg.noise.type != UVS_NOISE_NONE);
    const bool traced = io_traced(a.io);
    constexpr bool use_D = METHOD != UVS_METHOD_KF;          // Joseph with R_hat != R (GMCKF) / c != 1 (IMCC-KF)
    constexpr int FB = 32 / M;                               // frames of noise drawn at once by a covariance warp
    constexpr int LOGW = 6 + 2 * M;
    // =================================================================== chain warp: group g = filter g, lane gl = row
    const int gl = lane & 7, g = lane >> 3;
    WS& w = QUAD_WS(M, g);
    QuadCtl& ctl = ctlv[g];
    const bool rowl = gl < M;                                            // this lane carries a measurement row
    const double desired = rowl ? cfg.desired_f[gl] : 0.0;
    const double* target = cfg.targets + 3 * ((rowl ? gl : 0) >> 1);
    const bool servo = (cfg.v_max > 0.0) || (cfg.a_max > 0.0) || (cfg.servo_gain > 0.0 && cfg.servo_gain != 1.0);
    for (long long b0 = (long long)blockIdx.x * kQuadFilters; b0 < a.B; b0 += (long long)gridDim.x * kQuadFilters) {
        const long long b = b0 + g;
        const bool trial_ok = b < a.B;
        const long long bb = trial_ok ? b : 0;
        const uint64_t trial = a.trial0 + (uint64_t)bb * a.stride;
        double f_l = 0.0, q_l = 0.0;
        {
            double q[6];
            if (a.io.q0_in) {
#pragma unroll
                for (int j = 0; j < 6; ++j) q[j] = a.io.q0_in[bb * 6 + j];
            } else start_pose(cfg, trial, q);
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (gl == j) { q_l = q[j]; w.dq[j] = 0.0; }
            // ---- initial state (experiment.py:70-117)
            UR10 kin;
            Tf T6;
            ur10_fk(q, T6, &kin);
            double zc = 1.0, u = 0.0, v = 0.0, depth;
            if (rowl) {
                u = project_coord(T6, target, 0, depth);
                v = project_coord(T6, target, 1, depth);
                zc = camera_distance(T6, target);
            }
            if (a.io.X0_in) {
                if (cfg.initial_guess && rowl) f_l = (gl & 1) ? v : u;
            } else if (cfg.initial_guess) {
                double Jg[6][6];
                ur10_jacobian(kin, Jg);
                if (rowl) {
                    double row[6];
                    initial_guess_row(kin, Jg, gl, u, v, zc, row);
#pragma unroll
                    for (int j = 0; j < 6; ++j) w.X[6 * gl + j] = row[j];
                    f_l = (gl & 1) ? v : u;
                }
            } else {
                for (int r = gl; r < N; r += 8) {                        // experiment.py:117
                    double ua, ub;
                    uniforms(cfg.noise.seed, trial, 0, (uint32_t)r, 0, PURPOSE_X0, ua, ub);
                    w.X[r] = ua;
                }
            }
        }
        HoldState hs;
        double ise = 0.0, iae = 0.0, itae = 0.0;
        int status = UVS_STATUS_SUCCESS, frames = 0, flags_all = 0;
        double v_prev = 0.0;
        bool alive = trial_ok;

        // plant of frame k (joint angles -> this lane's noise-free feature coordinate); ends the trial when a feature
        // leaves the image (cfg.fov_check)
        double f_true = 0.0;
        auto plant = [&]() {
            if (gl < 6) {
                double sn, cs;
                sincos_bounded(dh_theta(gl, q_l), &sn, &cs);
                *reinterpret_cast<double2*>(w.sc + 2 * gl) = make_double2(sn, cs);
            }
            __syncwarp();
            double depth = 1.0;
            if (rowl) {
                double sn[6], cs[6], pc[3];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double2 t = *reinterpret_cast<const double2*>(w.sc + 2 * j);
                    sn[j] = t.x; cs[j] = t.y;
                }
                ur10_point_in_camera(sn, cs, target, pc);
                f_true = kPrincipal + fast_div(kFocal * ((gl & 1) ? pc[1] : pc[0]), pc[2]);    // pixel_of()
                depth = pc[2];
            }
            if (cfg.fov_check) {
                const bool bad = rowl && !(f_true >= 0.0 && f_true < kResolution && depth > kMinDepth);
                if (quad_group_any(bad, lane) && alive) {                // NaN centroid -> FAIL
                    alive = false; status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_OUT_OF_VIEW;
                }
            }
            __syncwarp();
        };
        auto publish_alive = [&]() {
            if (gl == 0) ctl.alive = alive ? 1 : 0;
            const bool any = __any_sync(UVS_FULL_MASK, alive);
            if (lane == 0) *quad_alive = any ? 1 : 0;
        };
        plant();
        publish_alive();
        quad_bar();                                                      // init
        QCLK_DECL(10)

        for (int k = 0; k < T; ++k) {
            // ---- noise, measurement, innovation, kernel weights ------------------------------------
            double nz = 0.0;
            if (a.io.noise_in) nz = (rowl && trial_ok) ? a.io.noise_in[((size_t)b * T + k) * M + gl] : 0.0;
            else if (philox) {
                const double fresh = rowl ? ctl.noise[(k / FB) & 1][(k % FB) * M + gl] : 0.0;      // drawn by the covariance warp
                nz = next_noise_prefetched(hs, cfg.noise.hold_cnt, fresh, gl, M);
            }
            const double f_new = f_true + nz;                            // experiment.py:130-135
            const double z = f_new - f_l;                                // experiment.py:170-177
            f_l = f_new;
            double sigma, inv_s2;
            if (sig_tab) { const double2 t = sig_tab[k]; sigma = t.x; inv_s2 = t.y; }
            else { sigma = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T); inv_s2 = 1.0 / (sigma * sigma); }
            double nu_l = 0.0;
            if (rowl) {
                double h[6], x[6];
                load6(w.dq, h);
                load6(w.X + 6 * gl, x);
                nu_l = z;
#pragma unroll
                for (int r = 0; r < 6; ++r) nu_l -= h[r] * x[r];
                w.nu[gl] = nu_l;
            }
            double kappa = 1.0, s0_scale = 1.0, w_lane = 1.0;
            if (METHOD == UVS_METHOD_GMCKF) {                            // experiment.py:274-276, 280
                const double c = rowl ? exp(-0.5 * (nu_l * nu_l) * inv_s2) : 1.0;
                kappa = c;
                w_lane = fast_div(1.0, c + 1e-6);
                if (rowl) w.Lc[8 + gl] = w_lane - 1.0;                   // D = R_hat - R (diagonal)
            } else if (METHOD == UVS_METHOD_IMCCKF) {                    // experiment.py:258-261
                double qq = 0.0;
#pragma unroll
                for (int j = 0; j < M; ++j) { const double ej = __shfl_sync(UVS_FULL_MASK, nu_l, j, 8); qq += ej * ej; }
                const double c = gaussian_kernel(sqrt(qq), sigma);
                s0_scale = c;
                if (rowl) w.Lc[8 + gl] = 1.0 - c;                        // D = (1 - c) R
            }
            // right-hand side of the IBVS law (experiment.py:306-310), for the covariance warp's J^T b
            const double b_l = kappa * (f_l - desired);
            if (rowl) w.ib[27 + gl] = b_l;
            QCLK(0);
            quad_wait<1>();                                              // (1) A, S0 ready
            QCLK(1);
            // ---- S = s0_scale S0 + R_hat, L D L^T (redundant in the 8 lanes of the group) ---------------
            bool ok;
            {
                double L[M * (M + 1) / 2];
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double wi = (METHOD == UVS_METHOD_GMCKF) ? __shfl_sync(UVS_FULL_MASK, w_lane, i, 8) : 1.0;
#pragma unroll
                    for (int j = 0; j < M; ++j)
                        if (j <= i) L[tri(i, j)] = s0_scale * w.S[i * M + j] + ((i == j) ? wi : 0.0);
                }
                ok = ldl_packed<M>(L);
                __syncwarp();                                            // every lane has read S0
                if (rowl) {                                              // column gl of S^-1 (symmetric: = row gl)
                    double e[M];
#pragma unroll
                    for (int j = 0; j < M; ++j) e[j] = (j == gl) ? 1.0 : 0.0;
                    ldl_solve<M>(L, e);
#pragma unroll
                    for (int j = 0; j < M; j += 2) *reinterpret_cast<double2*>(w.S + gl * M + j) = make_double2(e[j], e[j + 1]);
                    // y = gain_scale S^-1 nu: the state correction is X += A y (covariance warp, while this warp forms K)
                    double y = 0.0;
#pragma unroll
                    for (int j = 0; j < M; ++j) y += e[j] * w.nu[j];
                    w.Lc[gl] = ((METHOD == UVS_METHOD_IMCCKF) ? s0_scale : 1.0) * y;
                }
                if (gl == 0) ctl.ok = ok ? 1 : 0;
            }
            QCLK(2);
            quad_arrive<2>();                                            // (2) S^-1, y published (the group's S^-1 is complete:
            QCLK(3);                                                     //     the arrival starts with a __syncwarp)
            if (ok && alive) quad_gain_rows_chain<M, UVS_QUAD_GAIN_RB>(w, (METHOD == UVS_METHOD_IMCCKF) ? s0_scale : 1.0, use_D, gl);
            QCLK(4);
            quad_wait<3>();                                              // (3) K, U | X, Gram matrix, J^T b published
            QCLK(5);
            // ---- IBVS: dq = -gain pinv(J) (kappa o error)                  experiment.py:300-312
            double x[6];
            int ib_flags = 0;
            quad_pinv_from_gram<M>(w.X, b_l, w.ib, gl, x, ib_flags, w.jacobi_scratch());
            QCLK(6);
            double dq_l = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
                if (gl == r) dq_l = -cfg.ibvs_gain * x[r];
            const int fr_flags = (ok ? 0 : UVS_FLAG_NOT_SPD) | ib_flags;
            const bool bad = quad_group_any(!isfinite(dq_l), lane);     // np.linalg.pinv raises (experiment.py:313-316)
            if (alive) {
                flags_all |= fr_flags;
                if (bad) { alive = false; status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_NONFINITE; }
            }
            const bool done = alive;                                     // frame k completed
            const size_t bt = (size_t)b * T + k;
            double dq_cmd = dq_l;                                        // the command the plant executes
            if (done && gl < 6) {
                if (a.io.dq_in) dq_cmd = a.io.dq_in[bt * 6 + gl];        // replay gate: recorded command
                w.dq[gl] = dq_cmd;                                       // experiment.py:183-188: next frame's regressor
            }
            publish_alive();
            if (gl == 0) ctl.done = done ? 1 : 0;
            QCLK(7);
            // (4): the command is out — the covariance warps start the next frame's predict (when their rank update is
            // done: this warp does not wait for it) while this warp logs, integrates and evaluates the next frame's plant
            const bool any_alive = __any_sync(UVS_FULL_MASK, alive);
            quad_arrive<4>();
            QCLK(8);
            if (done) {
                // logging (experiment.py:323-331) and error integrals (plot_errorbar.m:39-84)
                if (a.io.log) {
                    double* lg = a.io.log + bt * LOGW;
                    if (gl < 6) lg[gl] = q_l;
                    if (rowl) { lg[6 + gl] = f_l; lg[6 + M + gl] = nz; }
                }
                if (rowl) {
                    const double t = (double)(k + 1) * cfg.dt;
                    const double e = desired - f_l;
                    ise += e * e; iae += fabs(e); itae += t * fabs(e);
                }
                ++frames;
                if (traced) {
                    if (a.io.dq_log && gl < 6) a.io.dq_log[bt * 6 + gl] = dq_l;
                    if (a.io.kappa_log && rowl) a.io.kappa_log[bt * M + gl] = kappa;
                    if (a.io.flags_log && gl == 0) a.io.flags_log[bt] = fr_flags;
                }
                if (gl < 6) {                                            // experiment.py:320,335
                    double v = dq_cmd;
                    if (servo) { v = servo_rate(cfg, dq_cmd, v_prev); v_prev = v; }
                    q_l = q_l + v * cfg.dt;
                }
            }
            if (!any_alive) break;
            if (k + 1 < T) {
                plant();                                                 // plant of the next frame
                // a feature left the image: the trial ends BEFORE that frame's predict (which the covariance warp has
                // started meanwhile; it restores P when it sees this after barrier (1))
                if (gl == 0 && !alive) ctl.alive = 0;
            }
            QCLK(9);
        }
        QCLK_FLUSH(10, 12);
        // ||.||_2 over the M coordinates
        const double s0 = sqrt(quad_group_sum(ise * ise)), s1 = sqrt(quad_group_sum(iae * iae)),
                     s2 = sqrt(quad_group_sum(itae * itae));
        if (trial_ok && gl == 0) {
            a.io.stats[b * 3 + 0] = s0;
            a.io.stats[b * 3 + 1] = s1;
            a.io.stats[b * 3 + 2] = s2;
            a.io.status[b] = status;
            if (a.io.frames) a.io.frames[b] = frames;
            if (a.io.flags) a.io.flags[b] = flags_all;
        }
        quad_bar();
    }
}

template <int M, int METHOD>
__global__ void UVS_QUAD_BOUNDS dense_quad_kernel(const __grid_constant__ ClosedLoopArgs a) {
    static_assert(M >= 6 && M <= 8 && METHOD != UVS_METHOD_MCKF, "quad kernel: M in {6, 8}, KF / IMCC-KF / GMCKF");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const int T = cfg.T;
    const size_t tab_bytes = sigma_table_bytes(T);
    double2* tab = reinterpret_cast<double2*>(smem_raw);
    const int warp = threadIdx.x >> 5;
    if (tab_bytes != 0) {
        for (int k = threadIdx.x; k < T; k += blockDim.x) {
            const double sg = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T);
            tab[k] = make_double2(sg, 1.0 / (sg * sg));
        }
    }
    __syncthreads();

    if (warp < kQuadFilters) quad_covariance_role<M, METHOD>(a, warp);
    else quad_chain_role<M, METHOD>(a);
}

// 1: FAST closed-loop launches of the dense family with M >= 6 and KF / IMCC-KF / GMCKF run the quad kernel
#ifndef UVS_DENSE_QUAD
#define UVS_DENSE_QUAD 1
#endif

template <int M, int METHOD>
static cudaError_t quad_launch(const ClosedLoopArgs& a, int sms, cudaStream_t st) {
    auto kernel = dense_quad_kernel<M, METHOD>;
    const size_t smem = quad_smem_bytes<M>(a.cfg.T);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kQuadThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (const char* lim = getenv("UVS_DENSE_CTAS_PER_SM")) {      // tuning experiments only
        const int v = atoi(lim);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    if (getenv("UVS_DEBUG_LAUNCH")) fprintf(stderr, "[uvs] dense_quad_kernel<%d,%d>: %d CTAs/SM, %zu B smem\n", M, METHOD, per_sm, smem);
    long long grid = (long long)sms * per_sm;
    const long long need = (a.B + kQuadFilters - 1) / kQuadFilters;
    if (need < grid) grid = need;
    kernel<<<(int)grid, kQuadThreads, smem, st>>>(a);
    return cudaGetLastError();
}

// does this launch qualify?  (the caller has already established the FAST conditions)
static inline bool quad_enabled() {
    if (!UVS_DENSE_QUAD) return false;
    const char* e = getenv("UVS_DENSE_QUAD");                     // UVS_DENSE_QUAD=0: one filter per warp (comparison runs)
    return !(e && e[0] == '0');
}

// defines launch_dense_quad<MM> (declared in dense_kernels.cuh) in dense_quad_m{6,8}.cu
#define UVS_DEFINE_QUAD_LAUNCHER(MM)                                                                                  \
    template <> bool launch_dense_quad<MM>(const ClosedLoopArgs& a, int sms, cudaStream_t st, cudaError_t* err) {      \
        if (!quad_enabled()) return false;                                                                             \
        switch (a.cfg.filt.method) {                                                                                   \
        case UVS_METHOD_KF: *err = quad_launch<MM, UVS_METHOD_KF>(a, sms, st); return true;                            \
        case UVS_METHOD_IMCCKF: *err = quad_launch<MM, UVS_METHOD_IMCCKF>(a, sms, st); return true;                    \
        case UVS_METHOD_GMCKF: *err = quad_launch<MM, UVS_METHOD_GMCKF>(a, sms, st); return true;                      \
        default: return false;                                                                                         \
        }                                                                                                              \
    }

}  // namespace uvs
