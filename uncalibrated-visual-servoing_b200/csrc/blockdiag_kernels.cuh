// blockdiag_kernels.cuh — the block-diagonal family: one measurement row per lane (kernels and per-M launchers; instantiated
// per measurement count in blockdiag_m{2,6,8}.cu, dispatched by blockdiag_kernels.cu).
//
// With the reference's hard-coded P0 = Q = I_N, R = I_m (experiment.py:73-76) and H = I_m (x) h^T
// (experiment.py:188) the covariance stays exactly block-diagonal (m blocks of 6x6): the N-state
// filter is m independent 6-state filters with one scalar measurement each, sharing the regressor
// h and coupled only through the MCKF stopping rule, the IMCC-KF scalar weight and the IBVS
// pseudo-inverse (SURVEY.md §7 "structure trap", Appendix C item 5).  Here lane g of a group of GS
// lanes owns row g: its 6 states and the packed lower triangle of its 6x6 block live in registers
// for the whole trial; a warp carries 32/GS trials.
//
// Per row (P^- = P + I, v = P^- h, s0 = h^T v, nu = z - h^T x, R = I):
//   KF      K = v / (s0 + 1)
//   GMCKF   K = v / (s0 + 1/(G(nu) + 1e-6))
//   IMCCKF  K = c v / (c s0 + 1),  c = G(||nu||_2 over the group)
//   MCKF    fixed point with Cx = G(L^-1 (x - xc)), Cy = G(z - h^T xc), L = chol(P^-)
//   P = P^- - K v^T - v K^T + (s0 + 1) K K^T            (Joseph, experiment.py:297)
//   P = P^- - K v^T                                      (joseph = 0, tests/kalman_1_feature.py:126)
#pragma once
#include "blockdiag_kernels.h"
#include <cstdlib>
#include "smallmat.cuh"

// resident CTAs per SM asked of ptxas for the closed-loop kernel (register cap 65536 / (MINB * 128))
#ifndef UVS_BD_MINB
#define UVS_BD_MINB 3
#endif

namespace uvs {

template <int GS>
UVS_DEV double group_sum(double v) {
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

template <int GS>
UVS_DEV bool group_any(bool p) {
    unsigned b = __ballot_sync(UVS_FULL_MASK, p);
    const int lane = threadIdx.x & 31;
    const unsigned gm = (GS == 32) ? 0xffffffffu : (((1u << GS) - 1u) << (lane & ~(GS - 1)));
    return (b & gm) != 0u;
}

struct RowState {
    double x[6];
    double P[21];      // packed lower triangle of this row's 6x6 covariance block
};

// one predict + update of this lane's row; `own` = lane owns a row of an existing trial
template <int GS, int METHOD>
UVS_DEV double bd_step(RowState& s, const uvs_filter_params& fp, const double (&h)[6], double z, double sigma,
                       bool own, int& flags, int& epoch_out) {
    // predict: P^- = P + I
#pragma unroll
    for (int i = 0; i < 6; ++i) s.P[tri(i, i)] += 1.0;
    double v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) a += s.P[i >= j ? tri(i, j) : tri(j, i)] * h[j];
        v[i] = a;
    }
    double s0 = 0.0, nu = z;
#pragma unroll
    for (int i = 0; i < 6; ++i) { s0 += h[i] * v[i]; nu -= h[i] * s.x[i]; }
    double K[6];
    double kappa = 1.0;
    bool skip = false;

    if (METHOD == UVS_METHOD_KF) {
        const double inv = 1.0 / (s0 + 1.0);
#pragma unroll
        for (int i = 0; i < 6; ++i) { K[i] = v[i] * inv; s.x[i] += K[i] * nu; }
    } else if (METHOD == UVS_METHOD_GMCKF) {
        const double c = gaussian_kernel(nu, sigma);
        kappa = c;
        const double inv = 1.0 / (s0 + 1.0 / (c + 1e-6));
#pragma unroll
        for (int i = 0; i < 6; ++i) { K[i] = v[i] * inv; s.x[i] += K[i] * nu; }
    } else if (METHOD == UVS_METHOD_IMCCKF) {
        const double q = group_sum<GS>(own ? nu * nu : 0.0);
        const double c = gaussian_kernel(sqrt(q), sigma);
        const double inv = c / (c * s0 + 1.0);
#pragma unroll
        for (int i = 0; i < 6; ++i) { K[i] = v[i] * inv; s.x[i] += K[i] * nu; }
    } else {   // MCKF, experiment.py:194-250
        double L[21];
#pragma unroll
        for (int i = 0; i < 21; ++i) L[i] = s.P[i];
        const bool spd = chol_packed<6>(L);            // reciprocal diagonal
        bool bad = group_any<GS>(own && !spd);
        if (bad) { flags |= UVS_FLAG_NOT_SPD; skip = true; }
        double xc[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { xc[i] = s.x[i]; K[i] = 0.0; }
        int epoch = 0;
        double difference = INFINITY;
        bool active = !skip;
        while (__any_sync(UVS_FULL_MASK, active && difference > fp.fpi_threshold && epoch < fp.fpi_epoch_max)) {
            const bool run = active && difference > fp.fpi_threshold && epoch < fp.fpi_epoch_max;
            // whitened residuals and their weights
            double ey = z, ex[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { ey -= h[i] * xc[i]; ex[i] = s.x[i] - xc[i]; }
            chol_forward<6>(L, ex);
            double icx[6];
            bool cx0 = false;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double cx = gaussian_kernel(ex[i], sigma);
                cx0 = cx0 || (cx == 0.0);
                icx[i] = 1.0 / cx;
            }
            const double cy = gaussian_kernel(ey, sigma);
            const bool g_cx0 = group_any<GS>(own && cx0);
            const bool g_cy0 = group_any<GS>(own && cy == 0.0);
            // vhat = L diag(1/cx) L^T h
            double w[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                double a = (1.0 / L[tri(j, j)]) * h[j];
#pragma unroll
                for (int i = 0; i < 6; ++i) if (i > j) a += L[tri(i, j)] * h[i];
                w[j] = a * icx[j];
            }
            double vh[6], sh = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double a = (1.0 / L[tri(i, i)]) * w[i];
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j < i) a += L[tri(i, j)] * w[j];
                vh[i] = a;
                sh += h[i] * a;
            }
            const double inv = 1.0 / (sh + 1.0 / cy);
            double kn[6], xn[6], num = 0.0, den = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                kn[i] = vh[i] * inv;
                xn[i] = s.x[i] + kn[i] * nu;                      // experiment.py:242 (innovation of X, not Xc)
                num += (xn[i] - xc[i]) * (xn[i] - xc[i]);
                den += xc[i] * xc[i];
            }
            num = group_sum<GS>(own ? num : 0.0);
            den = group_sum<GS>(own ? den : 0.0);
            if (run) {
                if (g_cx0) { flags |= UVS_FLAG_CX_SINGULAR; skip = true; active = false; }
                else if (g_cy0) { flags |= UVS_FLAG_SKIP_SINGULAR; skip = true; active = false; }
                else {
#pragma unroll
                    for (int i = 0; i < 6; ++i) { xc[i] = xn[i]; K[i] = kn[i]; }
                    difference = sqrt(num) / sqrt(den);            // experiment.py:244
                    ++epoch;
                    if (epoch == fp.fpi_epoch_max && fp.mckf_skip) { flags |= UVS_FLAG_SKIP_MAXEPOCH; skip = true; }
                }
            }
        }
        epoch_out = epoch;
        if (!skip) {
#pragma unroll
            for (int i = 0; i < 6; ++i) s.x[i] = xc[i];
        }
    }
    if (!skip) {
        if (fp.joseph) {
            const double sj = s0 + 1.0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j <= i)
                    s.P[tri(i, j)] = s.P[tri(i, j)] - K[i] * v[j] - v[i] * K[j] + sj * K[i] * K[j];
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) if (j <= i) s.P[tri(i, j)] -= K[i] * v[j];
        }
    }
    return kappa;
}

template <int M> struct GroupSize { static constexpr int value = (M == 2) ? 2 : 8; };

UVS_DEV void load_row_P(RowState& s, const double* P, int layout, int M, int row) {
    const int N = 6 * M;
    if (P == nullptr) {
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; ++j) if (j <= i) s.P[tri(i, j)] = (i == j) ? 1.0 : 0.0;
        return;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) if (j <= i)
            s.P[tri(i, j)] = (layout == UVS_P_DENSE) ? P[(size_t)(6 * row + i) * N + 6 * row + j]
                                                     : P[row * 36 + i * 6 + j];
}

UVS_DEV void store_row_P(const RowState& s, double* P, int layout, int M, int row) {
    const int N = 6 * M;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const double v = s.P[i >= j ? tri(i, j) : tri(j, i)];
            if (layout == UVS_P_DENSE) P[(size_t)(6 * row + i) * N + 6 * row + j] = v;
            else P[row * 36 + i * 6 + j] = v;
        }
}

template <int M, int METHOD>
__global__ void __launch_bounds__(128) bd_replay_kernel(const __grid_constant__ ReplayArgs a) {
    constexpr int GS = GroupSize<M>::value, GPW = 32 / GS, N = 6 * M;
    const int lane = threadIdx.x & 31, gl = lane % GS, grp = lane / GS;
    const long long warp_id = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const size_t psz = (a.p_layout == UVS_P_DENSE) ? (size_t)N * N : (size_t)M * 36;
    for (long long b0 = warp_id * GPW; b0 < a.B; b0 += n_warps * GPW) {
        const long long b = b0 + grp;
        const bool own = (b < a.B) && (gl < M);
        RowState s;
#pragma unroll
        for (int i = 0; i < 6; ++i) s.x[i] = own ? a.X[b * N + 6 * gl + i] : 0.0;
        load_row_P(s, own && a.P ? a.P + b * psz : nullptr, a.p_layout, M, gl);
        double f_prev = (own && !a.direct) ? a.f_init[b * M + gl] : 0.0;
        for (int k = 0; k < a.T; ++k) {
            double h[6], z = 0.0, sigma;
            const long long bb = (b < a.B) ? b : 0;
            if (a.direct) {
#pragma unroll
                for (int r = 0; r < 6; ++r) h[r] = a.h[bb * 6 + r];
                if (own) z = a.Z[b * M + gl];
                sigma = a.sigma;
            } else {
                if (own) {
                    const double fk = a.f[((size_t)b * a.T + k) * M + gl];
                    z = fk - f_prev;
                    f_prev = fk;
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) h[r] = (k == 0) ? 0.0 : a.dq[((size_t)bb * a.T + k - 1) * 6 + r];
                sigma = annealed_bw(a.fp.kernel_bw, a.fp.annealing, k, a.k_max);
            }
            int flags = 0, epoch = 0;
            const double kappa = bd_step<GS, METHOD>(s, a.fp, h, z, sigma, own, flags, epoch);
            if (own) {
                const size_t bt = (size_t)b * a.T + k;
                if (a.X_out) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) a.X_out[bt * N + 6 * gl + i] = s.x[i];
                }
                if (a.P_out) store_row_P(s, a.P_out + bt * psz, a.p_layout, M, gl);
                if (a.kappa_out) a.kappa_out[bt * M + gl] = kappa;
                if (a.flags_out && gl == 0) a.flags_out[bt] = flags | (epoch << UVS_FLAG_EPOCH_SHIFT);
            }
        }
        if (own) {
#pragma unroll
            for (int i = 0; i < 6; ++i) a.X[b * N + 6 * gl + i] = s.x[i];
            if (a.P) store_row_P(s, a.P + b * psz, a.p_layout, M, gl);
        }
    }
}

// EXTRAS: the joint servo limits (cfg.v_max / a_max / servo_gain), the per-frame trace and the dq_in replay gate
// (uvs_closed_loop_io, ABI 4) live in a separate instantiation, so that the production launch — kinematic plant, no trace —
// does not carry their registers through the frame loop (everything of this family is register-resident: the trace code
// alone took the kernel from 332 to 786 bytes of spill stores and cost 8-15 % of throughput when it was unconditional)
template <int M, int METHOD, bool EXTRAS>
__global__ void __launch_bounds__(128, UVS_BD_MINB) bd_closed_loop_kernel(const __grid_constant__ ClosedLoopArgs a) {
    constexpr int GS = GroupSize<M>::value, GPW = 32 / GS, N = 6 * M, LOGW = 6 + 2 * M;
    constexpr int WPC = 4;
    // per-group strides chosen so that the GPW groups of a warp start 8 banks apart (x 4-byte banks, mod 32): with the
    // natural strides (48 / 56 doubles = 0 / 16 banks mod 32) the groups' identical access patterns collided 2-4 ways
#ifndef UVS_BD_PAD
#define UVS_BD_PAD 1
#endif
    constexpr int JS = M * 6 + ((UVS_BD_PAD && M == 8) ? 4 : 0), SS = 56 + (UVS_BD_PAD ? 4 : 0);
    __shared__ __align__(16) double sJ[WPC][GPW][JS];
    __shared__ __align__(16) double sSc[WPC][GPW][SS];
    __shared__ __align__(16) double sBig[(M >= 6) ? WPC : 1][(M >= 6) ? GPW : 1][(M >= 6) ? pinv_scratch_doubles(M) : 2];   // Jacobi fallback work matrices
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gl = lane % GS, grp = lane / GS;
    const long long warp_id = (long long)blockIdx.x * WPC + warp;
    const long long n_warps = (long long)gridDim.x * WPC;
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const int T = cfg.T;
    double* Jg = sJ[warp][grp];
    double* sc = sSc[warp][grp];

    for (long long b0 = warp_id * GPW; b0 < a.B; b0 += n_warps * GPW) {
        const long long b = b0 + grp;
        const bool trial_ok = b < a.B;
        const bool own = trial_ok && (gl < M);
        const long long bb = trial_ok ? b : 0;
        const uint64_t trial = a.trial0 + (uint64_t)bb * a.stride;
        double rho = cfg.noise.rho, alpha = cfg.noise.alpha;
        if (cfg.sweep_len > 0 && cfg.sweep_param != 0) {
            long long lvl = (long long)(trial / (uint64_t)(cfg.sweep_epoch > 0 ? cfg.sweep_epoch : 1));
            if (lvl >= cfg.sweep_len) lvl = cfg.sweep_len - 1;
            if (cfg.sweep_param == 1) rho = cfg.sweep[lvl]; else alpha = cfg.sweep[lvl];
        }
        double q[6];
        if (a.io.q0_in) {
#pragma unroll
            for (int j = 0; j < 6; ++j) q[j] = a.io.q0_in[bb * 6 + j];
        } else {
#pragma unroll
            for (int j = 0; j < 6; ++j) q[j] = cfg.q_start[j];
            if (cfg.change_q_start) {
                double r0, r1, d;
                uniforms(cfg.pose_seed, trial, 0, 0, 0, PURPOSE_POSE, r0, d);
                uniforms(cfg.pose_seed, trial, 0, 1, 0, PURPOSE_POSE, r1, d);
                q[0] = cfg.q_start[0] + 2 * (r0 - 1) * (M_PI / 18);
                q[1] = cfg.q_start[1] + 2 * (r1 - 1) * (M_PI / 9);
            }
        }
        RowState s;
        double f_l = 0.0;
        const int row = (gl < M) ? gl : 0;
        {
            UR10 kin;
            Tf T6;
            ur10_fk(q, T6, &kin);
            const double* pw = cfg.targets + 3 * (row >> 1);
            double depth;
            double u = project_coord(T6, pw, 0, depth);
            double v = project_coord(T6, pw, 1, depth);
            const double zc = camera_distance(T6, pw);               // computeZ: DH-6 origin to disc centre, whatever the plant
            if (EXTRAS && scene_plant(cfg)) {                            // what the detector measures at the start pose
                double sn0[6], cs0[6];
                bool vis;
#pragma unroll
                for (int j = 0; j < 6; ++j) sincos(dh_theta(j, q[j]), &sn0[j], &cs0[j]);
                u = scene_feature(sn0, cs0, pw, cfg, 0, depth, vis);
                v = scene_feature(sn0, cs0, pw, cfg, 1, depth, vis);
            }
            if (a.io.X0_in) {
#pragma unroll
                for (int i = 0; i < 6; ++i) s.x[i] = a.io.X0_in[bb * N + 6 * row + i];
                if (cfg.initial_guess) f_l = (row & 1) ? v : u;
            } else if (cfg.initial_guess) {
                double Jgeo[6][6];
                ur10_jacobian(kin, Jgeo);
                initial_guess_row(kin, Jgeo, row, u, v, zc, s.x);
                f_l = (row & 1) ? v : u;
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    double ua, ub;
                    uniforms(cfg.noise.seed, trial, 0, (uint32_t)(6 * row + i), 0, PURPOSE_X0, ua, ub);
                    s.x[i] = ua;
                }
            }
        }
        load_row_P(s, nullptr, UVS_P_BLOCKS, M, row);
        HoldState hs;
        double ise = 0.0, iae = 0.0, itae = 0.0;
        double dq[6] = {0, 0, 0, 0, 0, 0};
        constexpr bool servo = EXTRAS;
        double v_prev[6] = {0, 0, 0, 0, 0, 0};      // joint rates executed in the previous frame (EXTRAS only)
        int status = UVS_STATUS_SUCCESS, frames = 0, flags_all = 0;
        bool alive = trial_ok;
        const double desired = cfg.desired_f[row];
        const bool traced = EXTRAS && (a.io.X_log != nullptr || a.io.P_log != nullptr || a.io.dq_log != nullptr ||
                                       a.io.kappa_log != nullptr || a.io.flags_log != nullptr);

        for (int k = 0; k < T; ++k) {
            if (!__any_sync(UVS_FULL_MASK, alive)) break;
            const double t = (double)(k + 1) * cfg.dt;
            // plant: sin/cos of the DH angles split over the group's lanes, then this lane's target point walked
            // back through the chain into the camera frame (ur10_simulation.py:97-110 + Appendix A pixel model)
            double sn[6], cs[6], pc[3];
            ur10_sincos_group<GS>(q, gl, sn, cs);
            double f_true, depth;
            bool visible;
            if (EXTRAS && scene_plant(cfg)) {
                f_true = scene_feature(sn, cs, cfg.targets + 3 * (row >> 1), cfg, row & 1, depth, visible);
            } else {
                ur10_point_in_camera(sn, cs, cfg.targets + 3 * (row >> 1), pc);
                f_true = pixel_of(pc, row & 1);
                depth = pc[2];
                visible = (f_true >= 0.0 && f_true < kResolution && depth > kMinDepth);
            }
            if (cfg.fov_check) {
                const bool bad = own && !visible;
                if (group_any<GS>(bad) && alive) { alive = false; status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_OUT_OF_VIEW; }
            }
            double nz;
            if (a.io.noise_in) nz = own ? a.io.noise_in[((size_t)b * T + k) * M + gl] : 0.0;
            else if (cfg.noise.type == UVS_NOISE_NONE) nz = 0.0;
            else nz = next_noise(hs, cfg.noise, rho, alpha, trial, (uint32_t)k, own ? gl : 64, M);
            const double f_new = f_true + nz;
            const double z = f_new - f_l;
            double h[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) h[r] = (k == 0) ? 0.0 : dq[r];
            const double sigma = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T);
            int flags = 0, epoch = 0;
            const double kappa = bd_step<GS, METHOD>(s, cfg.filt, h, z, sigma, own && alive, flags, epoch);
            if (alive && (flags & kAbortFlags)) {                // the reference ends FAIL here (common.cuh)
                alive = false; status = UVS_STATUS_FAIL; flags_all |= flags;
            }
            // IBVS on the group's Jacobian
            __syncwarp();
            if (gl < M) {
#pragma unroll
                for (int i = 0; i < 6; ++i) Jg[gl * 6 + i] = s.x[i];
            }
            __syncwarp();
            double x[6];
            int ib_flags = 0;
            pinv_group<M, GS>(Jg, kappa * (f_new - desired), sc, gl, x, ib_flags, (M >= 6) ? sBig[(M >= 6) ? warp : 0][(M >= 6) ? grp : 0] : nullptr);
            bool bad = false;
            double dq_new[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) { dq_new[r] = -cfg.ibvs_gain * x[r]; bad = bad || !isfinite(dq_new[r]); }
#pragma unroll
            for (int i = 0; i < 6; ++i) bad = bad || (own && !isfinite(s.x[i]));
            bad = group_any<GS>(bad);
            if (alive && bad) { alive = false; status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_NONFINITE; }
            if (alive) {
                f_l = f_new;
                flags_all |= flags | ib_flags;
                if (a.io.log && trial_ok) {
                    double* lg = a.io.log + ((size_t)b * T + k) * LOGW;
                    if (gl == 0) {
#pragma unroll
                        for (int r = 0; r < 6; ++r) lg[r] = q[r];
                    }
                    if (own) { lg[6 + gl] = f_l; lg[6 + M + gl] = nz; }
                }
                if (own) {
                    const double e = desired - f_l;
                    ise += e * e; iae += fabs(e); itae += t * fabs(e);
                }
                ++frames;
                if (EXTRAS && traced && trial_ok) {                      // per-frame trace (uvs_closed_loop_io, ABI 4)
                    const size_t bt = (size_t)b * T + k;
                    if (own) {
                        if (a.io.X_log) {
#pragma unroll
                            for (int i = 0; i < 6; ++i) a.io.X_log[bt * N + 6 * gl + i] = s.x[i];
                        }
                        if (a.io.P_log) store_row_P(s, a.io.P_log + bt * (M * 36), UVS_P_BLOCKS, M, gl);
                        if (a.io.kappa_log) a.io.kappa_log[bt * M + gl] = kappa;
                    }
                    if (gl == 0) {
                        if (a.io.dq_log) {
#pragma unroll
                            for (int r = 0; r < 6; ++r) a.io.dq_log[bt * 6 + r] = dq_new[r];
                        }
                        if (a.io.flags_log) a.io.flags_log[bt] = flags | ib_flags | (epoch << UVS_FLAG_EPOCH_SHIFT);
                    }
                }
                if (EXTRAS && a.io.dq_in && trial_ok) {                  // replay gate: recorded command
#pragma unroll
                    for (int r = 0; r < 6; ++r) dq_new[r] = a.io.dq_in[((size_t)b * T + k) * 6 + r];
                }
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    dq[r] = dq_new[r];
                    if (EXTRAS && cfg.servo_substeps > 0) {
                        servo_substepped(cfg, dq[r], q[r], v_prev[r]);
                    } else {
                        double v = dq[r];
                        if (servo) { v = servo_rate(cfg, dq[r], v_prev[r]); v_prev[r] = v; }
                        q[r] += v * cfg.dt;
                    }
                }
            }
        }
        const double s0 = sqrt(group_sum<GS>(own ? ise * ise : 0.0));
        const double s1 = sqrt(group_sum<GS>(own ? iae * iae : 0.0));
        const double s2 = sqrt(group_sum<GS>(own ? itae * itae : 0.0));
        if (trial_ok && gl == 0) {
            a.io.stats[b * 3 + 0] = s0;
            a.io.stats[b * 3 + 1] = s1;
            a.io.stats[b * 3 + 2] = s2;
            a.io.status[b] = status;
            if (a.io.frames) a.io.frames[b] = frames;
            if (a.io.flags) a.io.flags[b] = flags_all;
        }
        if (own) {
            if (a.io.X_out) {
#pragma unroll
                for (int i = 0; i < 6; ++i) a.io.X_out[b * N + 6 * gl + i] = s.x[i];
            }
            if (a.io.P_out) {
                const size_t psz = (a.io.p_layout == UVS_P_DENSE) ? (size_t)N * N : (size_t)M * 36;
                store_row_P(s, a.io.P_out + b * psz, a.io.p_layout, M, gl);
            }
        }
        __syncwarp();
    }
}

// grid = whole waves of resident CTAs (occupancy API), so that the statically strided trial loop ends evenly on every SM
#ifndef UVS_BD_WAVES
#define UVS_BD_WAVES 16
#endif
template <typename Kernel>
static long long bd_grid(Kernel kernel, long long need, int sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    int waves = UVS_BD_WAVES;
    if (const char* e = getenv("UVS_BD_WAVES")) {          // tuning experiments only
        const int v = atoi(e);
        if (v >= 1) waves = v;
    }
    const long long grid = (long long)sms * per_sm * waves;
    return need < grid ? need : grid;
}

template <int M, int METHOD>
static cudaError_t bd_replay_launch(const ReplayArgs& a, int sms, cudaStream_t st) {
    constexpr int GPW = 32 / GroupSize<M>::value;
    const long long warps = (a.B + GPW - 1) / GPW;
    const long long grid = bd_grid(bd_replay_kernel<M, METHOD>, (warps + 3) / 4, sms);
    bd_replay_kernel<M, METHOD><<<(int)grid, 128, 0, st>>>(a);
    return cudaGetLastError();
}

template <int M, int METHOD>
static cudaError_t bd_closed_launch(const ClosedLoopArgs& a, int sms, cudaStream_t st) {
    constexpr int GPW = 32 / GroupSize<M>::value;
    const long long warps = (a.B + GPW - 1) / GPW;
    const long long need = (warps + 3) / 4;
    const uvs_closed_loop_io& io = a.io;
    const bool extras = (a.cfg.v_max > 0.0) || (a.cfg.a_max > 0.0) || (a.cfg.servo_gain > 0.0 && a.cfg.servo_gain != 1.0) ||
                        io.dq_in || io.X_log || io.P_log || io.dq_log || io.kappa_log || io.flags_log || scene_plant(a.cfg);
    if (extras)
        bd_closed_loop_kernel<M, METHOD, true><<<(int)bd_grid(bd_closed_loop_kernel<M, METHOD, true>, need, sms), 128, 0, st>>>(a);
    else
        bd_closed_loop_kernel<M, METHOD, false><<<(int)bd_grid(bd_closed_loop_kernel<M, METHOD, false>, need, sms), 128, 0, st>>>(a);
    return cudaGetLastError();
}

template <int M>
cudaError_t bd_replay_m(const ReplayArgs& a, int sms, cudaStream_t st) {
    switch (a.fp.method) {
    case UVS_METHOD_KF: return bd_replay_launch<M, UVS_METHOD_KF>(a, sms, st);
    case UVS_METHOD_MCKF: return bd_replay_launch<M, UVS_METHOD_MCKF>(a, sms, st);
    case UVS_METHOD_IMCCKF: return bd_replay_launch<M, UVS_METHOD_IMCCKF>(a, sms, st);
    case UVS_METHOD_GMCKF: return bd_replay_launch<M, UVS_METHOD_GMCKF>(a, sms, st);
    default: return cudaErrorInvalidValue;
    }
}

template <int M>
cudaError_t bd_closed_m(const ClosedLoopArgs& a, int sms, cudaStream_t st) {
    switch (a.cfg.filt.method) {
    case UVS_METHOD_KF: return bd_closed_launch<M, UVS_METHOD_KF>(a, sms, st);
    case UVS_METHOD_MCKF: return bd_closed_launch<M, UVS_METHOD_MCKF>(a, sms, st);
    case UVS_METHOD_IMCCKF: return bd_closed_launch<M, UVS_METHOD_IMCCKF>(a, sms, st);
    case UVS_METHOD_GMCKF: return bd_closed_launch<M, UVS_METHOD_GMCKF>(a, sms, st);
    default: return cudaErrorInvalidValue;
    }
}

#define UVS_DEFINE_BLOCKDIAG_LAUNCHERS(MM)                                                      \
    template cudaError_t bd_replay_m<MM>(const ReplayArgs& a, int sms, cudaStream_t st);       \
    template cudaError_t bd_closed_m<MM>(const ClosedLoopArgs& a, int sms, cudaStream_t st);

}  // namespace uvs
