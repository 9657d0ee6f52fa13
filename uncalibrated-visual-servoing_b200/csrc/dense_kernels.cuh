// dense_kernels.cuh — kernels of the dense (general SPD covariance) family: one filter per warp.
//   dense_replay_kernel       uvs_filter_step / uvs_replay  (experiment.py:166-300, open loop)
//   dense_closed_loop_kernel  uvs_closed_loop               (experiment.py:48-359, whole trial)
// Instantiated per measurement count M in dense_m{2,6,8}.cu.
#pragma once
#include "dense.cuh"
#include "kernel_args.cuh"
#include <cstdlib>

// CTA geometry: kDenseWarps filters per CTA, UVS_DENSE_MINB CTAs per SM asked of ptxas (register cap
// 65536 / (MINB * 128)); the launchers size the persistent grid with the occupancy API.
#ifndef UVS_DENSE_MINB
#define UVS_DENSE_MINB 2
#endif
// 1: closed-loop launches with Q = R = NULL and the Joseph form use the instantiation that knows it at compile time
#ifndef UVS_DENSE_FAST
#define UVS_DENSE_FAST 1
#endif

#ifdef UVS_DENSE_MAXNREG
#define UVS_DENSE_BOUNDS __maxnreg__(UVS_DENSE_MAXNREG)
#else
#define UVS_DENSE_BOUNDS __launch_bounds__(kDenseThreads, UVS_DENSE_MINB)
#endif

namespace uvs {

#ifndef UVS_DENSE_WARPS
#define UVS_DENSE_WARPS 4
#endif
constexpr int kDenseWarps = UVS_DENSE_WARPS;
constexpr int kDenseThreads = kDenseWarps * 32;

template <int M, bool MCKF>
__host__ __device__ constexpr size_t dense_smem_bytes(int warps) {
    return sizeof(DenseConsts<M>) + (size_t)warps * sizeof(DenseWS<M, MCKF>);
}

// closed loop: per-CTA table {sigma_k, 1 / sigma_k^2} of the annealed kernel bandwidth (experiment.py:196-199), filled
// once per CTA with the IEEE division instead of once per trial and frame; trials longer than this fall back to it
constexpr int kSigmaTableMax = 1024;
__host__ __device__ constexpr size_t sigma_table_bytes(int T) {
    return (T <= kSigmaTableMax) ? (size_t)T * 2 * sizeof(double) : 0;
}

template <int M, int METHOD>
__global__ void UVS_DENSE_BOUNDS dense_replay_kernel(const __grid_constant__ ReplayArgs a) {
    constexpr int N = 6 * M;
    constexpr bool MCKF = METHOD == UVS_METHOD_MCKF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseConsts<M>& cc = *reinterpret_cast<DenseConsts<M>*>(smem_raw);
    DenseWS<M, MCKF>* wsv = reinterpret_cast<DenseWS<M, MCKF>*>(smem_raw + sizeof(DenseConsts<M>));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    if (threadIdx.x == 0) init_consts<M>(cc, a.fp);
    __syncthreads();
    DenseWS<M, MCKF>& w = wsv[warp];
    const size_t psz = (a.p_layout == UVS_P_DENSE) ? (size_t)N * N : (size_t)M * 36;

    for (long long b = (long long)blockIdx.x * wpc + warp; b < a.B; b += (long long)gridDim.x * wpc) {
        for (int r = lane; r < N; r += 32) w.X[r] = a.X[b * N + r];
        PBlock pb;
        load_P<M>(pb, w.Pd, a.P ? a.P + b * psz : nullptr, a.p_layout, lane);
        __syncwarp();
        double f_prev = (!a.direct && lane < M) ? a.f_init[b * M + lane] : 0.0;
        for (int k = 0; k < a.T; ++k) {
            double h[6], z = 0.0, sigma;
            if (a.direct) {
#pragma unroll
                for (int r = 0; r < 6; ++r) h[r] = a.h[b * 6 + r];
                if (lane < M) z = a.Z[b * M + lane];
                sigma = a.sigma;
            } else {
                if (lane < M) {
                    const double fk = a.f[((size_t)b * a.T + k) * M + lane];
                    z = fk - f_prev;                                    // experiment.py:170-177
                    f_prev = fk;
                }
#pragma unroll
                for (int r = 0; r < 6; ++r)                             // experiment.py:183-188
                    h[r] = (k == 0) ? 0.0 : a.dq[((size_t)b * a.T + k - 1) * 6 + r];
                sigma = annealed_bw(a.fp.kernel_bw, a.fp.annealing, k, a.k_max);
            }
            StepOut so;
#ifdef UVS_PHASE_CLOCKS
            unsigned long long clk_acc[UVS_NPHASE] = {};
            long long clk_last = 0;
#endif
            const double kappa = dense_step<M, METHOD>(pb, w, cc, a.fp, h, z, sigma, lane, so, 0.0 UVS_CLK_PASS);
            const size_t bt = (size_t)b * a.T + k;
            if (a.X_out) for (int r = lane; r < N; r += 32) a.X_out[bt * N + r] = w.X[r];
            if (a.P_out) store_P<M>(pb, w.Pd, a.P_out + bt * psz, a.p_layout, lane);
            if (a.kappa_out && lane < M) a.kappa_out[bt * M + lane] = kappa;
            if (a.flags_out && lane == 0) a.flags_out[bt] = so.flags | (so.epoch << UVS_FLAG_EPOCH_SHIFT);
            if (a.K_out) for (int e = lane; e < N * M; e += 32) a.K_out[bt * N * M + e] = w.Kt[(e % M) * (N + 2) + e / M];
            __syncwarp();
        }
        for (int r = lane; r < N; r += 32) a.X[b * N + r] = w.X[r];
        if (a.P) store_P<M>(pb, w.Pd, a.P + b * psz, a.p_layout, lane);
        __syncwarp();
    }
}

// FAST: Q = R = NULL (identity) and the Joseph update, known at compile time (dense_step<..., FAST>).
template <int M, int METHOD, bool FAST>
__global__ void UVS_DENSE_BOUNDS dense_closed_loop_kernel(const __grid_constant__ ClosedLoopArgs a) {
    constexpr int N = 6 * M;
    constexpr bool MCKF = METHOD == UVS_METHOD_MCKF;
    constexpr int LOGW = 6 + 2 * M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseConsts<M>& cc = *reinterpret_cast<DenseConsts<M>*>(smem_raw);
    DenseWS<M, MCKF>* wsv = reinterpret_cast<DenseWS<M, MCKF>*>(smem_raw + sizeof(DenseConsts<M>));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const int T = cfg.T;
    const double2* sig_tab = nullptr;
    if (sigma_table_bytes(T) != 0) {
        double2* tab = reinterpret_cast<double2*>(smem_raw + sizeof(DenseConsts<M>) + (size_t)wpc * sizeof(DenseWS<M, MCKF>));
        for (int k = threadIdx.x; k < T; k += blockDim.x) {
            const double sg = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T);
            tab[k] = make_double2(sg, 1.0 / (sg * sg));
        }
        sig_tab = tab;
    }
    if (threadIdx.x == 0) init_consts<M>(cc, cfg.filt);
    __syncthreads();
    DenseWS<M, MCKF>& w = wsv[warp];

    for (long long b = (long long)blockIdx.x * wpc + warp; b < a.B; b += (long long)gridDim.x * wpc) {
        const uint64_t trial = a.trial0 + (uint64_t)b * a.stride;
        // noise level of this trial (main.py:121-127)
        double rho = cfg.noise.rho, alpha = cfg.noise.alpha;
        if (cfg.sweep_len > 0 && cfg.sweep_param != 0) {
            long long lvl = (long long)(trial / (uint64_t)(cfg.sweep_epoch > 0 ? cfg.sweep_epoch : 1));
            if (lvl >= cfg.sweep_len) lvl = cfg.sweep_len - 1;
            if (cfg.sweep_param == 1) rho = cfg.sweep[lvl]; else alpha = cfg.sweep[lvl];
        }
        double f_l = 0.0;              // this lane's feature coordinate (noisy, as logged)
        {
            double q[6];
            if (a.io.q0_in) {
#pragma unroll
                for (int j = 0; j < 6; ++j) q[j] = a.io.q0_in[b * 6 + j];
            } else start_pose(cfg, trial, q);
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (lane == j) { w.q[j] = q[j]; w.dq[j] = 0.0; }

            // ---- initial state (experiment.py:70-117) -----------------------------------------
            UR10 kin;
            Tf T6;
            ur10_fk(q, T6, &kin);
            double zc = 1.0, u = 0.0, v = 0.0, depth;
            if (lane < M) {
                const double* pw = cfg.targets + 3 * (lane >> 1);
                u = project_coord(T6, pw, 0, depth);
                v = project_coord(T6, pw, 1, depth);
                zc = camera_distance(T6, pw);                       // computeZ: DH-6 origin to disc centre, whatever the plant
                if constexpr (!FAST) {
                    if (scene_plant(cfg)) {                             // what the detector measures at the start pose
                        double sn0[6], cs0[6];
                        bool vis;
#pragma unroll
                        for (int j = 0; j < 6; ++j) sincos(dh_theta(j, q[j]), &sn0[j], &cs0[j]);
                        u = scene_feature(sn0, cs0, pw, cfg, 0, depth, vis);
                        v = scene_feature(sn0, cs0, pw, cfg, 1, depth, vis);
                    }
                }
            }
            if (a.io.X0_in) {
                for (int r = lane; r < N; r += 32) w.X[r] = a.io.X0_in[b * N + r];
                if (cfg.initial_guess && lane < M) f_l = (lane & 1) ? v : u;
            } else if (cfg.initial_guess) {
                double Jg[6][6];
                ur10_jacobian(kin, Jg);
                if (lane < M) {
                    double row[6];
                    initial_guess_row(kin, Jg, lane, u, v, zc, row);
#pragma unroll
                    for (int j = 0; j < 6; ++j) w.X[6 * lane + j] = row[j];
                    f_l = (lane & 1) ? v : u;
                }
            } else {
                for (int r = lane; r < N; r += 32) {                    // experiment.py:117
                    double ua, ub;
                    uniforms(cfg.noise.seed, trial, 0, (uint32_t)r, 0, PURPOSE_X0, ua, ub);
                    w.X[r] = ua;
                }
            }
        }
        PBlock pb;
        load_P<M>(pb, w.Pd, cfg.P0, UVS_P_DENSE, lane);
        __syncwarp();

        HoldState hs;
        double ise = 0.0, iae = 0.0, itae = 0.0;
        int status = UVS_STATUS_SUCCESS, frames = 0, flags_all = 0;
        const double desired = (lane < M) ? cfg.desired_f[lane] : 0.0;
        const double* target = cfg.targets + 3 * ((lane < M ? lane : 0) >> 1);
        const bool servo = (cfg.v_max > 0.0) || (cfg.a_max > 0.0) || (cfg.servo_gain > 0.0 && cfg.servo_gain != 1.0);
        double v_prev = 0.0;               // lane j < 6: joint rate executed in the previous frame (servo limits only)

        constexpr int FB = 32 / M;         // frames of noise drawn at once: lane = slot * M + coordinate
        const bool philox = (a.io.noise_in == nullptr) && (cfg.noise.type != UVS_NOISE_NONE);
        const bool traced = io_traced(a.io);
        double fresh_batch = 0.0;
#ifdef UVS_PHASE_CLOCKS
        unsigned long long clk_acc[UVS_NPHASE] = {};
        long long clk_last = clock64();
#endif
        for (int k = 0; k < T; ++k) {
            UVS_CLK(11);
            // ---- plant: joint angles -> this lane's feature coordinate ------------------------------
            // lane j < 6 evaluates sin/cos of DH angle j; lanes < M walk their target point back through
            // the chain into the camera frame (ur10_simulation.py:97-110 + SURVEY Appendix A pixel model)
            double q_l = 0.0;
            if (lane < 6) {
                q_l = w.q[lane];
                double sn, cs;
                sincos_bounded(dh_theta(lane, q_l), &sn, &cs);
                *reinterpret_cast<double2*>(w.sc + 2 * lane) = make_double2(sn, cs);
            }
            __syncwarp();
            double f_true = 0.0, depth = 1.0;
            [[maybe_unused]] bool visible = true;      // general instantiation only (scene plant); FAST tests the bounds directly
            if (lane < M) {
                double sn[6], cs[6], pc[3];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double2 t = *reinterpret_cast<const double2*>(w.sc + 2 * j);
                    sn[j] = t.x; cs[j] = t.y;
                }
                if constexpr (FAST) {
                    ur10_point_in_camera(sn, cs, target, pc);
                    f_true = kPrincipal + fast_div(kFocal * ((lane & 1) ? pc[1] : pc[0]), pc[2]);    // pixel_of()
                    depth = pc[2];
                } else if (scene_plant(cfg)) {
                    f_true = scene_feature(sn, cs, target, cfg, lane & 1, depth, visible);
                } else {
                    ur10_point_in_camera(sn, cs, target, pc);
                    f_true = kPrincipal + fast_div(kFocal * ((lane & 1) ? pc[1] : pc[0]), pc[2]);    // pixel_of()
                    depth = pc[2];
                    visible = (f_true >= 0.0 && f_true < kResolution && depth > kMinDepth);
                }
            }
            if (cfg.fov_check) {
                bool bad;
                if constexpr (FAST) bad = (lane < M) && !(f_true >= 0.0 && f_true < kResolution && depth > kMinDepth);
                else bad = (lane < M) && !visible;
                if (__any_sync(UVS_FULL_MASK, bad)) {                   // NaN centroid -> FAIL
                    status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_OUT_OF_VIEW; break;
                }
            }
            UVS_CLK(0);
            double nz = 0.0;
            if (a.io.noise_in) nz = (lane < M) ? a.io.noise_in[((size_t)b * T + k) * M + lane] : 0.0;
            else if (philox) {
                const int slot = k % FB;
                if (slot == 0) {
                    const int fs = lane / M;
                    if (fs < FB && k + fs < T)
                        fresh_batch = noise_sample(cfg.noise, rho, alpha, trial, (uint32_t)(k + fs), (uint32_t)(lane % M));
                }
                const double fresh = shfl_d(fresh_batch, slot * M + (lane % M));
                nz = next_noise_prefetched(hs, cfg.noise.hold_cnt, fresh, lane, M);
            }
            UVS_CLK(1);
            const double f_new = f_true + nz;                            // experiment.py:130-135
            const double z = f_new - f_l;                                // experiment.py:170-177
            f_l = f_new;
            double sigma, inv_s2;
            if (sig_tab) { const double2 t = sig_tab[k]; sigma = t.x; inv_s2 = t.y; }
            else { sigma = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T); inv_s2 = 1.0 / (sigma * sigma); }
            StepOut so;
            double kappa;
            {
                double h[6];                                             // experiment.py:183-188: the previous
                load6(w.dq, h);                                          // command (zeros on the first frame)
                kappa = dense_step<M, METHOD, (M >= 6 && !MCKF), FAST>(pb, w, cc, cfg.filt, h, z, sigma, lane, so, inv_s2 UVS_CLK_PASS);
            }
            flags_all |= so.flags;
            if (so.flags & kAbortFlags) { status = UVS_STATUS_FAIL; break; }      // the reference ends FAIL here (common.cuh)

            // IBVS: dq = -gain pinv(J) (kappa o error)                    experiment.py:300-312
            double x[6];
            int ib_flags = 0;
            if constexpr (M >= 6 && !MCKF) {
                if (so.rank_pending)
                    ibvs_with_rank_update<M>(pb, w.Pd, w.Kt, w.Ut, w.X, kappa * (f_l - desired), w.ib, lane, x, ib_flags, w.jacobi_scratch() UVS_CLK_PASS);
                else
                    pinv_warp<M>(w.X, kappa * (f_l - desired), w.ib, lane, x, ib_flags, w.jacobi_scratch());
            } else {
                pinv_warp<M>(w.X, kappa * (f_l - desired), w.ib, lane, x, ib_flags, w.jacobi_scratch());
            }
            flags_all |= ib_flags;
            UVS_CLK(10);
            double dq_l = 0.0;
#pragma unroll
            for (int r = 0; r < 6; ++r)
                if (lane == r) dq_l = -cfg.ibvs_gain * x[r];
            // np.linalg.pinv raises on a non-finite Jacobian (experiment.py:313-316).  Testing the command is enough: a NaN or
            // Inf anywhere in X = J poisons every entry of the Gram matrix J^T J (each sums over all rows), so the
            // factorisation is rejected, the rank-revealing path returns NaN and dq is non-finite; no scan of X needed.
            const bool bad = !isfinite(dq_l);
            if (__any_sync(UVS_FULL_MASK, bad)) {
                status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_NONFINITE; break;
            }
            // logging (experiment.py:323-331) and error integrals (plot_errorbar.m:39-84)
            if (a.io.log) {
                double* lg = a.io.log + ((size_t)b * T + k) * LOGW;
                if (lane < 6) lg[lane] = q_l;
                if (lane < M) { lg[6 + lane] = f_l; lg[6 + M + lane] = nz; }
            }
            if (lane < M) {
                const double t = (double)(k + 1) * cfg.dt;
                const double e = desired - f_l;
                ise += e * e; iae += fabs(e); itae += t * fabs(e);
            }
            ++frames;
            if (traced) trace_frame<M>(a.io, (size_t)b * T + k, w.Pd, w.X, dq_l, kappa,
                                       so.flags | ib_flags | (so.epoch << UVS_FLAG_EPOCH_SHIFT), lane);
            if (a.io.dq_in && lane < 6) dq_l = a.io.dq_in[((size_t)b * T + k) * 6 + lane];     // replay gate: recorded command
            if (lane < 6) {                                              // experiment.py:320,335
                w.dq[lane] = dq_l;
                if constexpr (FAST) {
                    double v = dq_l;
                    if (servo) { v = servo_rate(cfg, dq_l, v_prev); v_prev = v; }
                    w.q[lane] = q_l + v * cfg.dt;
                } else if (cfg.servo_substeps > 0) {
                    double qn = q_l;
                    servo_substepped(cfg, dq_l, qn, v_prev);
                    w.q[lane] = qn;
                } else {
                    double v = dq_l;
                    if (servo) { v = servo_rate(cfg, dq_l, v_prev); v_prev = v; }
                    w.q[lane] = q_l + v * cfg.dt;
                }
            }
            __syncwarp();
        }
#ifdef UVS_PHASE_CLOCKS
        if (lane == 0)
            for (int i = 0; i < UVS_NPHASE; ++i) atomicAdd(a.phase_clk + i, clk_acc[i]);
#endif
        // ||.||_2 over the M coordinates
        double s0 = ise * ise, s1 = iae * iae, s2 = itae * itae;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s0 += shfl_xor_d(s0, o); s1 += shfl_xor_d(s1, o); s2 += shfl_xor_d(s2, o); }
        if (lane == 0) {
            a.io.stats[b * 3 + 0] = sqrt(s0);
            a.io.stats[b * 3 + 1] = sqrt(s1);
            a.io.stats[b * 3 + 2] = sqrt(s2);
            a.io.status[b] = status;
            if (a.io.frames) a.io.frames[b] = frames;
            if (a.io.flags) a.io.flags[b] = flags_all;
        }
        if (a.io.X_out) for (int r = lane; r < N; r += 32) a.io.X_out[b * N + r] = w.X[r];
        if (a.io.P_out) {
            const size_t psz = (a.io.p_layout == UVS_P_DENSE) ? (size_t)N * N : (size_t)M * 36;
            store_P<M>(pb, w.Pd, a.io.P_out + b * psz, a.io.p_layout, lane);
        }
        __syncwarp();
    }
}

// FAST closed-loop launch of the dense family through the warp-specialised quad kernel (dense_quad.cuh; its kernels are
// compiled in their own translation units dense_quad_m{6,8}.cu): true (and *err set) when it took the launch.
template <int M> bool launch_dense_quad(const ClosedLoopArgs& a, int sms, cudaStream_t st, cudaError_t* err);
template <> inline bool launch_dense_quad<2>(const ClosedLoopArgs&, int, cudaStream_t, cudaError_t*) { return false; }
template <> bool launch_dense_quad<6>(const ClosedLoopArgs& a, int sms, cudaStream_t st, cudaError_t* err);
template <> bool launch_dense_quad<8>(const ClosedLoopArgs& a, int sms, cudaStream_t st, cudaError_t* err);

// host-side dispatch (defined once per M in dense_m*.cu): persistent grid = SMs x resident CTAs
template <int M> cudaError_t launch_dense_replay(const ReplayArgs& a, int sms, cudaStream_t st);
template <int M> cudaError_t launch_dense_closed_loop(const ClosedLoopArgs& a, int sms, cudaStream_t st);

template <typename Kernel, typename Args>
static cudaError_t dense_launch(Kernel kernel, const Args& a, long long B, size_t smem, int sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kDenseThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (const char* lim = getenv("UVS_DENSE_CTAS_PER_SM")) {      // tuning experiments only
        const int v = atoi(lim);
        if (v >= 1 && v < per_sm) per_sm = v;
    }
    long long grid = (long long)sms * per_sm;
    const long long need = (B + kDenseWarps - 1) / kDenseWarps;
    if (need < grid) grid = need;
    kernel<<<(int)grid, kDenseThreads, smem, st>>>(a);
    return cudaGetLastError();
}

#define UVS_DEFINE_DENSE_LAUNCHERS(MM)                                                                       \
    template <int METHOD> static cudaError_t replay_one_##MM(const ReplayArgs& a, int sms, cudaStream_t st) { \
        return dense_launch(dense_replay_kernel<MM, METHOD>, a, a.B,                                         \
                            dense_smem_bytes<MM, METHOD == UVS_METHOD_MCKF>(kDenseWarps), sms, st);          \
    }                                                                                                        \
    template <int METHOD> static cudaError_t closed_one_##MM(const ClosedLoopArgs& a, int sms, cudaStream_t st) { \
        const uvs_filter_params& fp = a.cfg.filt;                                                            \
        if (UVS_DENSE_FAST && fp.Q == nullptr && fp.R == nullptr && fp.joseph && !scene_plant(a.cfg)) {      \
            cudaError_t qe;                                                                                  \
            if (METHOD != UVS_METHOD_MCKF && launch_dense_quad<MM>(a, sms, st, &qe)) return qe;   /* four filters per CTA */ \
            return dense_launch(dense_closed_loop_kernel<MM, METHOD, UVS_DENSE_FAST != 0>, a, a.B,           \
                                dense_smem_bytes<MM, METHOD == UVS_METHOD_MCKF>(kDenseWarps) +               \
                                    sigma_table_bytes(a.cfg.T), sms, st);                                    \
        }                                                                                                    \
        return dense_launch(dense_closed_loop_kernel<MM, METHOD, false>, a, a.B,                             \
                            dense_smem_bytes<MM, METHOD == UVS_METHOD_MCKF>(kDenseWarps) +                   \
                                sigma_table_bytes(a.cfg.T), sms, st);                                        \
    }                                                                                                        \
    template <> cudaError_t launch_dense_replay<MM>(const ReplayArgs& a, int sms, cudaStream_t st) {         \
        switch (a.fp.method) {                                                                               \
        case UVS_METHOD_KF: return replay_one_##MM<UVS_METHOD_KF>(a, sms, st);                               \
        case UVS_METHOD_MCKF: return replay_one_##MM<UVS_METHOD_MCKF>(a, sms, st);                           \
        case UVS_METHOD_IMCCKF: return replay_one_##MM<UVS_METHOD_IMCCKF>(a, sms, st);                       \
        case UVS_METHOD_GMCKF: return replay_one_##MM<UVS_METHOD_GMCKF>(a, sms, st);                         \
        default: return cudaErrorInvalidValue;                                                               \
        }                                                                                                    \
    }                                                                                                        \
    template <> cudaError_t launch_dense_closed_loop<MM>(const ClosedLoopArgs& a, int sms, cudaStream_t st) { \
        switch (a.cfg.filt.method) {                                                                         \
        case UVS_METHOD_KF: return closed_one_##MM<UVS_METHOD_KF>(a, sms, st);                               \
        case UVS_METHOD_MCKF: return closed_one_##MM<UVS_METHOD_MCKF>(a, sms, st);                           \
        case UVS_METHOD_IMCCKF: return closed_one_##MM<UVS_METHOD_IMCCKF>(a, sms, st);                       \
        case UVS_METHOD_GMCKF: return closed_one_##MM<UVS_METHOD_GMCKF>(a, sms, st);                         \
        default: return cudaErrorInvalidValue;                                                               \
        }                                                                                                    \
    }

}  // namespace uvs
