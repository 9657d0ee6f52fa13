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
//               and does the WIDE work of the frame: P += Q, A = P H^T, S0 = H A, the gain rows K = A S^-1 (two
//               interleaved substitutions per lane, the factor read from shared memory), X += K nu, U, P -= K U^T; and,
//               in the shadow of the chain warp's pseudo-inverse, the Philox / alpha-stable draws of the next 32 / M frames;
//   warp 4      "chain warp": group g = lanes 8g..8g+7 runs the SCALAR chains of filter g, one measurement row per lane,
//               the four filters of the CTA in lockstep: plant (sincos, point in camera, pixel), hold logic, innovation,
//               kernel weights, S = S0 + R_hat and its L D L^T, the IBVS pseudo-inverse (pinv_group with 8-lane groups),
//               command, error integrals, logging — one instruction stream for four filters instead of four.
// Neither role needs more than ~136 registers, so THREE CTAs = 12 filters are resident per SM (the shared-memory work
// space of a filter is unchanged, 14.5 KB), and the chain instructions are issued once per four filters.
//
// The roles exchange through the filter's shared-memory work space (DenseWS, as before) and meet at four CTA-wide named
// barriers per frame:
//   (4 / init) dq, alive, noise published | covariance: predict, A, S0        chain: noise, innovation, weights
//   (1)        A, S0, nu, weights ready   |                                    chain: S, L D L^T -> Lc, ok
//   (2)        factor ready               | covariance: gain rows, X, U
//   (3)        X ready                    | covariance: P -= K U^T (+ noise)  chain: pinv, dq, integrals, next plant
// Arithmetic: every quantity is computed by the same expression, in the same order, as in dense_step<M, METHOD, DEFER,
// FAST> + ibvs_with_rank_update (dense.cuh), so the 1e-9 replay gate of tests/test_gpu_parity.py applies unchanged.
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
constexpr int kQuadThreads = (kQuadFilters + 1) * 32;

struct alignas(16) QuadCtl {           // per filter
    double gain_scale;                 // IMCC-KF: c (1 otherwise)
    int ok;                            // S factorised (else: NOT_SPD, the correction is skipped)
    int alive;                         // the trial is still running
};

// the CTA's dynamic shared memory as seen by either role (declared inside each function so that every access is an LDS /
// STS on the shared window, not a generic load through a pointer argument)
#define QUAD_SMEM_VIEW(MM)                                                                                        \
    extern __shared__ __align__(16) unsigned char smem_raw[];                                                      \
    DenseWS<MM, false>* wsv = reinterpret_cast<DenseWS<MM, false>*>(smem_raw + sigma_table_bytes(a.cfg.T));        \
    QuadCtl* ctlv = reinterpret_cast<QuadCtl*>(smem_raw + sigma_table_bytes(a.cfg.T) +                             \
                                               kQuadFilters * sizeof(DenseWS<MM, false>));                         \
    volatile int* quad_alive = reinterpret_cast<volatile int*>(ctlv + kQuadFilters);

UVS_DEV void quad_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kQuadThreads) : "memory"); }

UVS_DEV bool quad_group_any(bool p, int lane) {
    const unsigned b = __ballot_sync(UVS_FULL_MASK, p);
    return ((b >> (lane & ~7)) & 0xffu) != 0u;
}

UVS_DEV double quad_group_sum(double v) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

template <int M>
__host__ __device__ constexpr size_t quad_smem_bytes(int T) {
    return sigma_table_bytes(T) + kQuadFilters * (sizeof(DenseWS<M, false>) + sizeof(QuadCtl)) + 16;
}

// Gain rows of one filter by its covariance warp: K = gain_scale A S^-1, X += K nu, U = A + (A S^-1) diag(ddiag) (use_D).
// S^-1 is the explicit symmetric inverse the chain warp left in w.S (the reference also forms it: K = P H^T inv(S),
// experiment.py:192, 286-287): a row of K is then eight independent 8-term dot products instead of two dependent
// substitutions, and with the j loop ROLLED the phase needs 16 + ~12 registers next to the warp's 72 of P — unrolled
// substitutions let ptxas hoist the factor's loads and push half of P into local memory (seen in the SASS).
template <int M>
UVS_DEV void quad_gain_rows(DenseWS<M, false>& w, double gain_scale, bool use_D, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    const double* ddiag = w.ib + 8;
#pragma unroll 1
    for (int row = lane; row < N; row += 32) {
        double kx[M];
#pragma unroll
        for (int k = 0; k < M; ++k) kx[k] = 0.0;
#pragma unroll 1
        for (int j = 0; j < M; ++j) {
            const double aj = w.At[j * LD + row];
            const double2* sr = reinterpret_cast<const double2*>(w.S + j * M);      // row j of S^-1 (broadcast)
#pragma unroll
            for (int k = 0; k < M; k += 2) {
                const double2 t = sr[k / 2];
                kx[k] += aj * t.x;
                kx[k + 1] += aj * t.y;
            }
        }
        double dxr = 0.0;
#pragma unroll
        for (int k = 0; k < M; ++k) {
            w.Kt[k * LD + row] = gain_scale * kx[k];
            dxr += kx[k] * w.nu[k];
        }
        w.X[row] += gain_scale * dxr;                            // experiment.py:193 / 265 / 291
        if (use_D) {
#pragma unroll
            for (int k = 0; k < M; ++k) w.Ut[k * LD + row] = w.At[k * LD + row] + kx[k] * ddiag[k];
        }
    }
    __syncwarp();
}

// P -= K U^T on the lower blocks by the covariance warp: rank_update() of dense.cuh with the k loop of the off-diagonal
// part ROLLED.  This warp is not on the frame's critical path (the chain warp is), so it trades the instruction-level
// parallelism of the unrolled form for registers: unrolled, ptxas hoists the loads of several k above the FMAs and parks
// P itself in local memory for the whole frame loop (measured in the SASS: 72 LDL.64 + 36 STL.64 per frame).
template <int M>
UVS_DEV void quad_rank_update(PBlock& pb, double* Pd, const double* Kt, const double* Ut, int lane) {
    constexpr int N = 6 * M, LD = N + 2;
    const double* kp = Kt + 6 * pb.bi;
    const double* up = Ut + 6 * pb.bj;
#pragma unroll 1
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
    // =================================================================== covariance warp of filter `warp`
    WS& w = wsv[warp];
    QuadCtl& ctl = ctlv[warp];
    volatile int* v_alive = &ctl.alive;
    volatile int* v_ok = &ctl.ok;
    for (long long b0 = (long long)blockIdx.x * kQuadFilters; b0 < a.B; b0 += (long long)gridDim.x * kQuadFilters) {
        const long long b = b0 + warp;
        const bool trial_ok = b < a.B;
        const long long bb = trial_ok ? b : 0;
        const uint64_t trial = a.trial0 + (uint64_t)bb * a.stride;
        if (a.io.X0_in)
            for (int r = lane; r < N; r += 32) w.X[r] = a.io.X0_in[bb * N + r];
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
            const bool live = *v_alive != 0;
            if (live) {
                double h[6];                                             // experiment.py:183-188: the previous command
                load6(w.dq, h);
                predict_and_A<M>(pb, w.Pd, w.At, h, nullptr, 1, ln);
                __syncwarp();
                for (int e = ln; e < M * M; e += 32) {                 // S0 = H A (build_and_factor_S)
                    const int i = e / M, kk = e % M;
                    double av[6];
                    load6(w.At + kk * LD + 6 * i, av);
                    double s0 = 0.0;
#pragma unroll
                    for (int r = 0; r < 6; ++r) s0 += h[r] * av[r];
                    w.S[e] = s0;
                }
            }
            QCLK(0);
            quad_bar();                                                  // (1)
            quad_bar();                                                  // (2) factor published
            QCLK(1);
            const bool ok = live && (*v_ok != 0);
            if (ok) quad_gain_rows<M>(w, ctl.gain_scale, use_D, ln);
            QCLK(2);
            quad_bar();                                                  // (3) X published
            QCLK(3);
            if (ok) {
                quad_rank_update<M>(pb, w.Pd, w.Kt, use_D ? w.Ut : w.At, ln);
                __syncwarp();
            }
            QCLK(4);
            quad_bar();                                                  // (4) dq, alive published
            QCLK(5);
            if (traced && live && (*v_alive != 0) && trial_ok) {         // state / covariance after the update of frame k
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
    constexpr int LOGW = 6 + 2 * M;
    // =================================================================== chain warp: group g = filter g, lane gl = row
    const int gl = lane & 7, g = lane >> 3;
    WS& w = wsv[g];
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
        double rho = cfg.noise.rho, alpha = cfg.noise.alpha;             // noise level of this trial (main.py:121-127)
        if (cfg.sweep_len > 0 && cfg.sweep_param != 0) {
            long long lvl = (long long)(trial / (uint64_t)(cfg.sweep_epoch > 0 ? cfg.sweep_epoch : 1));
            if (lvl >= cfg.sweep_len) lvl = cfg.sweep_len - 1;
            if (cfg.sweep_param == 1) rho = cfg.sweep[lvl]; else alpha = cfg.sweep[lvl];
        }
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
        // this lane's fresh draw of frame 0 (counter-based streams: a pure function of (trial, frame, coordinate))
        double fresh = (philox && rowl) ? noise_sample(cfg.noise, rho, alpha, trial, 0u, (uint32_t)gl) : 0.0;
        quad_bar();                                                      // init
        QCLK_DECL(10)

        for (int k = 0; k < T; ++k) {
            // ---- noise, measurement, innovation, kernel weights ------------------------------------
            double nz = 0.0;
            if (a.io.noise_in) nz = (rowl && trial_ok) ? a.io.noise_in[((size_t)b * T + k) * M + gl] : 0.0;
            else if (philox) {
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
                if (rowl) w.ib[8 + gl] = w_lane - 1.0;                   // D = R_hat - R (diagonal)
                if (gl == 0) ctl.gain_scale = 1.0;
            } else if (METHOD == UVS_METHOD_IMCCKF) {                    // experiment.py:258-261
                double qq = 0.0;
#pragma unroll
                for (int j = 0; j < M; ++j) { const double ej = __shfl_sync(UVS_FULL_MASK, nu_l, j, 8); qq += ej * ej; }
                const double c = gaussian_kernel(sqrt(qq), sigma);
                s0_scale = c;
                if (rowl) w.ib[8 + gl] = 1.0 - c;                        // D = (1 - c) R
                if (gl == 0) ctl.gain_scale = c;
            } else {
                if (gl == 0) ctl.gain_scale = 1.0;
            }
            QCLK(0);
            quad_bar();                                                  // (1) A, S0 ready
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
                }
                if (gl == 0) ctl.ok = ok ? 1 : 0;
            }
            QCLK(2);
            quad_bar();                                                  // (2)
            QCLK(3);
            // while the covariance warps solve for the gain rows: the fresh draw of the NEXT frame
            if (philox && rowl && k + 1 < T) fresh = noise_sample(cfg.noise, rho, alpha, trial, (uint32_t)(k + 1), (uint32_t)gl);
            QCLK(4);
            quad_bar();                                                  // (3) X published
            QCLK(5);
            // ---- IBVS: dq = -gain pinv(J) (kappa o error)                  experiment.py:300-312
            double x[6];
            int ib_flags = 0;
            pinv_group<M, 8>(w.X, kappa * (f_l - desired), w.ib, gl, x, ib_flags, w.jacobi_scratch());
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
            if (alive) {
                // logging (experiment.py:323-331) and error integrals (plot_errorbar.m:39-84)
                const size_t bt = (size_t)b * T + k;
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
                if (a.io.dq_in && gl < 6) dq_l = a.io.dq_in[bt * 6 + gl];                      // replay gate: recorded command
                if (gl < 6) {                                            // experiment.py:320,335
                    w.dq[gl] = dq_l;
                    double v = dq_l;
                    if (servo) { v = servo_rate(cfg, dq_l, v_prev); v_prev = v; }
                    q_l = q_l + v * cfg.dt;
                }
            }
            __syncwarp();
            QCLK(7);
            if (k + 1 < T) plant();                                      // plant of the next frame
            publish_alive();
            QCLK(8);
            quad_bar();                                                  // (4)
            QCLK(9);
            if (*quad_alive == 0) break;
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

// FAST closed-loop launch of the dense family: true (and *err set) when the quad kernel took it
template <int M, int METHOD>
static bool quad_try(const ClosedLoopArgs& a, int sms, cudaStream_t st, cudaError_t* err) {
    if constexpr (M >= 6 && METHOD != UVS_METHOD_MCKF) {
        if (!quad_enabled()) return false;
        *err = quad_launch<M, METHOD>(a, sms, st);
        return true;
    } else {
        return false;
    }
}

}  // namespace uvs
