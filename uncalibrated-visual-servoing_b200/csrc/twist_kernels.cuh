// twist_kernels.cuh — the camera-twist closed loop of the reference's tests/ scripts:
//   /root/reference/tests/{kalman,mckf}_{1,3,4}_feature(s).py   filter regressed on the commanded camera twist dp
//   /root/reference/tests/ibvs_{1,3,4}_feature(s).py            calibrated baseline (Method.ANALYTICAL)
// Per frame (kalman_4_features.py:111-197):
//   f = project(q) + noise ; Z = f - f_old ; H = kron(I_m, dp^T) ; filter update (dense.cuh, unchanged)
//   dp = -GAIN pinv(J) (f - f*) [; dp[i] = 0] ; dx = kron(I2, R06^T) dp ; dq = pinv(J_geom) dx ; q += dq TS
//   the loop stops once ||f - f*|| <= ERROR_THRESHOLD.
// One trial per warp, general covariance through dense_step<M, METHOD>; the joint state lives in registers
// (replicated on every lane), the geometric Jacobian is built column-per-lane while the DH chain is walked.
// Instantiated per measurement count in twist_m{2,6,8}.cu.
#pragma once
#include "dense.cuh"
#include "kernel_args.cuh"

namespace uvs {

constexpr int kTwistWarps = 4;
constexpr int kTwistThreads = kTwistWarps * 32;

struct TwistArgs {
    uvs_closed_loop_cfg cfg;
    uvs_twist_cfg tw;
    uvs_closed_loop_io io;
    unsigned long long trial0;
    long long B;
};

struct alignas(16) TwistWS {
    double Jg[36];          // geometric Jacobian, row-major (rows v;w, columns joints)
    double jac[72];         // work matrices of the cooperative pseudo-inverse fallback (wrist singularities)
    double R[9];            // camera rotation R06 used by the control law (the start pose's when tw.stale_fk)
    double pad;
};

template <int M, bool MCKF>
__host__ __device__ constexpr size_t twist_smem_bytes(int warps) {
    return sizeof(DenseConsts<M>) + (size_t)warps * (sizeof(DenseWS<M, MCKF>) + sizeof(TwistWS));
}

// One row of the analytic interaction matrix with raw pixel coordinates (kalman_4_features.py:86-99,
// ibvs_4_features.py:82-94).  coord = 2*feature + axis.
UVS_DEV void image_jacobian_row(int coord, double u, double v, double Zc, double f, double (&row)[6]) {
    if ((coord & 1) == 0) {
        row[0] = -f / Zc; row[1] = 0.0; row[2] = u / Zc; row[3] = u * v / f; row[4] = -(f * f + u * u) / f; row[5] = v;
    } else {
        row[0] = 0.0; row[1] = -f / Zc; row[2] = v / Zc; row[3] = (f * f + v * v) / f; row[4] = -u * v / f; row[5] = -u;
    }
}

// DH chain (ur10_simulation.py:97-110) that also leaves, on lane i < 6, the axis z_{i-1} and origin p_{i-1}
// of the frame joint i turns about (ur10_simulation.py:112-139) — the geometric Jacobian column of that lane.
UVS_DEV void ur10_chain_columns(const double (&s)[6], const double (&c)[6], int lane, Tf& T6, double (&z)[3],
                                double (&p)[3]) {
    Tf T;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) T.r[i][j] = (i == j) ? 1.0 : 0.0;
        T.p[i] = 0.0;
    }
    z[0] = 0.0; z[1] = 0.0; z[2] = 1.0;
    p[0] = p[1] = p[2] = 0.0;
#define UVS_KEEP(J)                                                              \
    if (lane == (J)) {                                                           \
        z[0] = T.r[0][2]; z[1] = T.r[1][2]; z[2] = T.r[2][2];                    \
        p[0] = T.p[0]; p[1] = T.p[1]; p[2] = T.p[2];                             \
    }
    dh_mul<0, -1>(T, s[0], c[0], 0.128, 0.0);    UVS_KEEP(1)
    dh_mul<1, 0>(T, s[1], c[1], 0.0, 0.6127);    UVS_KEEP(2)
    dh_mul<1, 0>(T, s[2], c[2], 0.0, 0.5716);    UVS_KEEP(3)
    dh_mul<0, -1>(T, s[3], c[3], 0.1639, 0.0);   UVS_KEEP(4)
    dh_mul<0, 1>(T, s[4], c[4], 0.1157, 0.0);    UVS_KEEP(5)
    dh_mul<1, 0>(T, s[5], c[5], 0.0922, 0.0);
#undef UVS_KEEP
    T6 = T;
}

// FAST: Q = R = NULL and the Joseph form known at compile time (as in dense_closed_loop_kernel); with M >= 6 the covariance
// update then runs inside the first pseudo-inverse (ibvs_with_rank_update).
template <int M, int METHOD, bool FAST>
__global__ void __launch_bounds__(kTwistThreads, 2) twist_closed_loop_kernel(const __grid_constant__ TwistArgs a) {
    constexpr int N = 6 * M;
    constexpr bool ANALYTIC = METHOD == UVS_METHOD_ANALYTICAL;
    constexpr bool MCKF = METHOD == UVS_METHOD_MCKF;
    constexpr int FILTER = ANALYTIC ? UVS_METHOD_KF : METHOD;      // ANALYTICAL never calls dense_step
    constexpr int LOGW = 6 + 2 * M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseConsts<M>& cc = *reinterpret_cast<DenseConsts<M>*>(smem_raw);
    DenseWS<M, MCKF>* wsv = reinterpret_cast<DenseWS<M, MCKF>*>(smem_raw + sizeof(DenseConsts<M>));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    TwistWS* twv = reinterpret_cast<TwistWS*>(smem_raw + sizeof(DenseConsts<M>) + (size_t)wpc * sizeof(DenseWS<M, MCKF>));
    const uvs_closed_loop_cfg& cfg = a.cfg;
    const uvs_twist_cfg& tw = a.tw;
    if (threadIdx.x == 0) init_consts<M>(cc, cfg.filt);
    __syncthreads();
    DenseWS<M, MCKF>& w = wsv[warp];
    TwistWS& ts = twv[warp];
    const int T = cfg.T;

    for (long long b = (long long)blockIdx.x * wpc + warp; b < a.B; b += (long long)gridDim.x * wpc) {
        const uint64_t trial = a.trial0 + (uint64_t)b * (uint64_t)(cfg.trial_stride > 1 ? cfg.trial_stride : 1);
        double rho = cfg.noise.rho, alpha = cfg.noise.alpha;
        if (cfg.sweep_len > 0 && cfg.sweep_param != 0) {                 // main.py:121-127
            long long lvl = (long long)(trial / (uint64_t)(cfg.sweep_epoch > 0 ? cfg.sweep_epoch : 1));
            if (lvl >= cfg.sweep_len) lvl = cfg.sweep_len - 1;
            if (cfg.sweep_param == 1) rho = cfg.sweep[lvl]; else alpha = cfg.sweep[lvl];
        }
        double q[6];                                                       // replicated on every lane
        if (a.io.q0_in) {
#pragma unroll
            for (int j = 0; j < 6; ++j) q[j] = a.io.q0_in[b * 6 + j];
        } else start_pose(cfg, trial, q);
        if (lane < 6) w.dq[lane] = 0.0;                                    // w.dq holds the last twist command dp
        const double* target = cfg.targets + 3 * ((lane < M ? lane : 0) >> 1);
        const double desired = (lane < M) ? cfg.desired_f[lane] : 0.0;

        // ---- initial feature vector and state (kalman_4_features.py:40,74-101; kalman_1:44; kalman_3:50) ----
        double f_l = 0.0;
        {
            Tf T6;
            ur10_fk(q, T6);
            double u = 0.0, v = 0.0, depth, zc = 1.0;
            if (lane < M) {
                u = project_coord(T6, target, 0, depth);
                v = project_coord(T6, target, 1, depth);
                zc = (tw.depth_mode == UVS_DEPTH_HEIGHT) ? T6.p[2] : camera_distance(T6, target);
                if (!FAST && scene_plant(cfg)) {
                    double sn0[6], cs0[6];
                    bool vis;
#pragma unroll
                    for (int j = 0; j < 6; ++j) sincos(dh_theta(j, q[j]), &sn0[j], &cs0[j]);
                    u = scene_feature(sn0, cs0, target, cfg, 0, depth, vis);
                    v = scene_feature(sn0, cs0, target, cfg, 1, depth, vis);
                }
            }
            if (tw.x0_mode == UVS_X0_IMAGE_JACOBIAN && lane < M) f_l = (lane & 1) ? v : u;
            if (a.io.X0_in) {
                for (int r = lane; r < N; r += 32) w.X[r] = a.io.X0_in[b * N + r];
            } else if (tw.x0_mode == UVS_X0_IMAGE_JACOBIAN) {
                if (lane < M) {
                    double row[6];
                    image_jacobian_row(lane, u, v, zc, tw.focal_x0, row);
#pragma unroll
                    for (int j = 0; j < 6; ++j) w.X[6 * lane + j] = row[j];
                }
            } else if (tw.x0_mode == UVS_X0_RANDOM) {
                for (int r = lane; r < N; r += 32) {
                    double ua, ub;
                    uniforms(cfg.noise.seed, trial, 0, (uint32_t)r, 0, PURPOSE_X0, ua, ub);
                    w.X[r] = ua;
                }
            } else {
                for (int r = lane; r < N; r += 32) w.X[r] = 0.0;
            }
        }
        PBlock pb;
        load_P<M>(pb, w.Pd, cfg.P0, UVS_P_DENSE, lane);
        __syncwarp();

        HoldState hs;
        double ise = 0.0, iae = 0.0, itae = 0.0;
        int status = UVS_STATUS_SUCCESS, frames = 0, flags_all = 0;
        const bool philox = (a.io.noise_in == nullptr) && (cfg.noise.type != UVS_NOISE_NONE);
        double err2 = (double)M;                                           // error = ones(m)  (kalman_4_features.py:70)
        const bool servo = (cfg.v_max > 0.0) || (cfg.a_max > 0.0) || (cfg.servo_gain > 0.0 && cfg.servo_gain != 1.0);
        double v_prev[6] = {0, 0, 0, 0, 0, 0};
        const bool traced = io_traced(a.io);
#ifdef UVS_PHASE_CLOCKS
        unsigned long long clk_acc[UVS_NPHASE] = {};
        long long clk_last = 0;
#endif

        for (int k = 0; k < T; ++k) {
            if (tw.err_threshold >= 0.0 && !(sqrt(err2) > tw.err_threshold)) break;     // kalman_4_features.py:111
            // ---- plant: DH chain, camera pose, this lane's Jacobian column and feature coordinate ----
            if (lane < 6) {
                double ql = q[0];
#pragma unroll
                for (int j = 1; j < 6; ++j) ql = (lane == j) ? q[j] : ql;
                double sn, cs;
                sincos(dh_theta(lane, ql), &sn, &cs);
                *reinterpret_cast<double2*>(w.sc + 2 * lane) = make_double2(sn, cs);
            }
            __syncwarp();
            Tf T6;
            double zax[3], org[3];
            double f_scene = 0.0, depth_scene = 1.0;
            bool vis_scene = true;
            bool scene = false;                                            // scene plant: general instantiation only
            if constexpr (!FAST) scene = scene_plant(cfg);
            {
                double sn[6], cs[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const double2 t = *reinterpret_cast<const double2*>(w.sc + 2 * j);
                    sn[j] = t.x; cs[j] = t.y;
                }
                ur10_chain_columns(sn, cs, lane, T6, zax, org);
                if constexpr (!FAST) {
                    if (scene && lane < M) f_scene = scene_feature(sn, cs, target, cfg, lane & 1, depth_scene, vis_scene);
                }
            }
            // tw.stale_fk: the scripts that never refresh UR10Simulation's cached transforms inside the loop keep the
            // start pose's J_geom and R06 (kalman_4_features.py:165-169; frame 0 is at the start pose)
            const bool refresh_fk = !tw.stale_fk || k == 0;
            if (refresh_fk && lane < 9) ts.R[lane] = T6.r[lane / 3][lane % 3];
            if (refresh_fk && lane < 6) {                                  // ur10_simulation.py:117-139
                const double dx = T6.p[0] - org[0], dy = T6.p[1] - org[1], dz = T6.p[2] - org[2];
                ts.Jg[0 * 6 + lane] = zax[1] * dz - zax[2] * dy;
                ts.Jg[1 * 6 + lane] = zax[2] * dx - zax[0] * dz;
                ts.Jg[2 * 6 + lane] = zax[0] * dy - zax[1] * dx;
                ts.Jg[3 * 6 + lane] = zax[0];
                ts.Jg[4 * 6 + lane] = zax[1];
                ts.Jg[5 * 6 + lane] = zax[2];
            }
            double f_true = 0.0, depth = 1.0, zc = 1.0;
            bool visible = true;
            if (lane < M) {
                f_true = project_coord(T6, target, lane & 1, depth);
                visible = (f_true >= 0.0 && f_true < kResolution && depth > kMinDepth);
                if (scene) { f_true = f_scene; depth = depth_scene; visible = vis_scene; }
                zc = (tw.depth_mode == UVS_DEPTH_HEIGHT) ? T6.p[2] : camera_distance(T6, target);
            }
            if (cfg.fov_check) {
                const bool bad = (lane < M) && !visible;
                if (__any_sync(UVS_FULL_MASK, bad)) {
                    status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_OUT_OF_VIEW; break;
                }
            }
            double nz = 0.0;
            if (a.io.noise_in) nz = (lane < M) ? a.io.noise_in[((size_t)b * T + k) * M + lane] : 0.0;
            else if (philox) nz = next_noise(hs, cfg.noise, rho, alpha, trial, (uint32_t)k, lane, M);
            const double f_new = f_true + nz;
            const double z = f_new - f_l;
            f_l = f_new;

            // ---- Jacobian: estimated (filter regressed on dp) or calibrated ----
            double kappa = 1.0;
            constexpr bool DEFER = FAST && M >= 6 && !MCKF && !ANALYTIC;
            bool rank_pending = false;
            int step_flags = 0;
            if constexpr (ANALYTIC) {
                const double f_p = shfl_d(f_l, lane ^ 1);
                if (lane < M) {
                    double row[6];
                    image_jacobian_row(lane, (lane & 1) ? f_p : f_l, (lane & 1) ? f_l : f_p, zc, tw.focal, row);
#pragma unroll
                    for (int j = 0; j < 6; ++j) w.X[6 * lane + j] = row[j];
                }
                __syncwarp();
            } else {
                double h[6];
                load6(w.dq, h);                                            // kalman_4_features.py:152-156
                const double sigma = annealed_bw(cfg.filt.kernel_bw, cfg.filt.annealing, k, T);
                StepOut so;
                kappa = dense_step<M, FILTER, DEFER, FAST>(pb, w, cc, cfg.filt, h, z, sigma, lane, so, 1.0 / (sigma * sigma) UVS_CLK_PASS);
                flags_all |= so.flags;
                step_flags = so.flags | (so.epoch << UVS_FLAG_EPOCH_SHIFT);
                rank_pending = so.rank_pending;
                if (so.flags & kAbortFlags) { status = UVS_STATUS_FAIL; break; }  // common.cuh: kAbortFlags
            }

            // ---- control: dp = -GAIN pinv(J) e ; dx = kron(I2, R^T) dp ; dq = pinv(J_geom) dx ----
            const double e_l = f_l - desired;
            double x[6];
            int ib_flags = 0;
            if constexpr (DEFER) {
                if (rank_pending) ibvs_with_rank_update<M>(pb, w.Pd, w.Kt, w.Ut, w.X, kappa * e_l, w.ib, lane, x, ib_flags, w.jacobi_scratch() UVS_CLK_PASS);
                else pinv_warp<M>(w.X, kappa * e_l, w.ib, lane, x, ib_flags, w.jacobi_scratch());
            } else {
                pinv_warp<M>(w.X, kappa * e_l, w.ib, lane, x, ib_flags, w.jacobi_scratch());
            }
            double dp[6];
#pragma unroll
            for (int r = 0; r < 6; ++r) dp[r] = ((tw.dp_zero_mask >> r) & 1) ? 0.0 : -cfg.ibvs_gain * x[r];
            double dx_l = 0.0;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double r0 = ts.R[c], r1 = ts.R[3 + c], r2 = ts.R[6 + c];
                const double lin = r0 * dp[0] + r1 * dp[1] + r2 * dp[2];
                const double ang = r0 * dp[3] + r1 * dp[4] + r2 * dp[5];
                dx_l = (lane == c) ? lin : ((lane == 3 + c) ? ang : dx_l);
            }
            double dq[6];
            int jg_flags = 0;
            pinv_warp<6>(ts.Jg, dx_l, w.ib, lane, dq, jg_flags, ts.jac);
            flags_all |= ib_flags | jg_flags;
            const int frame_flags = step_flags | ib_flags | jg_flags;

            bool bad = false;
#pragma unroll
            for (int r = 0; r < 6; ++r) bad = bad || !isfinite(dq[r]);
            for (int r = lane; r < N; r += 32) bad = bad || !isfinite(w.X[r]);
            if (__any_sync(UVS_FULL_MASK, bad)) {
                status = UVS_STATUS_FAIL; flags_all |= UVS_FLAG_NONFINITE; break;
            }
            if (a.io.log) {
                double* lg = a.io.log + ((size_t)b * T + k) * LOGW;
#pragma unroll
                for (int r = 0; r < 6; ++r)
                    if (lane == r) lg[r] = q[r];
                if (lane < M) { lg[6 + lane] = f_l; lg[6 + M + lane] = nz; }
            }
            double e2 = 0.0;
            if (lane < M) {
                const double t = (double)(k + 1) * cfg.dt;
                const double e = desired - f_l;
                ise += e * e; iae += fabs(e); itae += t * fabs(e);
                e2 = e * e;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e2 += shfl_xor_d(e2, o);
            err2 = e2;
            ++frames;
            if (traced) {                                                  // per-frame trace (uvs_closed_loop_io, ABI 4)
                double dq_l = dq[0];
#pragma unroll
                for (int r = 1; r < 6; ++r) dq_l = (lane == r) ? dq[r] : dq_l;
                trace_frame<M>(a.io, (size_t)b * T + k, w.Pd, w.X, dq_l, kappa, frame_flags, lane);
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                if (lane == r) w.dq[r] = dp[r];
                if (!FAST && cfg.servo_substeps > 0) {
                    servo_substepped(cfg, dq[r], q[r], v_prev[r]);
                } else {
                    double v = dq[r];
                    if (servo) { v = servo_rate(cfg, dq[r], v_prev[r]); v_prev[r] = v; }
                    q[r] += v * cfg.dt;                                    // kalman_4_features.py:173,192
                }
            }
            __syncwarp();
        }
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

template <int M> cudaError_t launch_twist_loop(const TwistArgs& a, int sms, cudaStream_t st);

template <typename Kernel>
static cudaError_t twist_launch(Kernel kernel, const TwistArgs& a, size_t smem, int sms, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTwistThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    long long grid = (long long)sms * per_sm;
    const long long need = (a.B + kTwistWarps - 1) / kTwistWarps;
    if (need < grid) grid = need;
    kernel<<<(int)grid, kTwistThreads, smem, st>>>(a);
    return cudaGetLastError();
}

#define UVS_DEFINE_TWIST_LAUNCHER(MM)                                                                            \
    template <int METHOD> static cudaError_t twist_one_##MM(const TwistArgs& a, int sms, cudaStream_t st) {     \
        const uvs_filter_params& fp = a.cfg.filt;                                                                \
        constexpr bool CAN_FAST = METHOD != UVS_METHOD_MCKF && METHOD != UVS_METHOD_ANALYTICAL;                  \
        if (CAN_FAST && fp.Q == nullptr && fp.R == nullptr && fp.joseph && !scene_plant(a.cfg))                  \
            return twist_launch(twist_closed_loop_kernel<MM, METHOD, CAN_FAST>, a,                               \
                                twist_smem_bytes<MM, METHOD == UVS_METHOD_MCKF>(kTwistWarps), sms, st);          \
        return twist_launch(twist_closed_loop_kernel<MM, METHOD, false>, a,                                      \
                            twist_smem_bytes<MM, METHOD == UVS_METHOD_MCKF>(kTwistWarps), sms, st);              \
    }                                                                                                            \
    template <> cudaError_t launch_twist_loop<MM>(const TwistArgs& a, int sms, cudaStream_t st) {                \
        switch (a.cfg.filt.method) {                                                                             \
        case UVS_METHOD_ANALYTICAL: return twist_one_##MM<UVS_METHOD_ANALYTICAL>(a, sms, st);                    \
        case UVS_METHOD_KF: return twist_one_##MM<UVS_METHOD_KF>(a, sms, st);                                    \
        case UVS_METHOD_MCKF: return twist_one_##MM<UVS_METHOD_MCKF>(a, sms, st);                                \
        case UVS_METHOD_IMCCKF: return twist_one_##MM<UVS_METHOD_IMCCKF>(a, sms, st);                            \
        case UVS_METHOD_GMCKF: return twist_one_##MM<UVS_METHOD_GMCKF>(a, sms, st);                              \
        default: return cudaErrorInvalidValue;                                                                   \
        }                                                                                                        \
    }

}  // namespace uvs
