// api.cu — the extern "C" boundary of libuvs_b200.so (declared in include/uvs_b200.h).
// Argument validation, kernel-family selection, launch geometry, error plumbing.
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "aux_kernels.h"
#include "blockdiag_kernels.h"
#include "dense_kernels.cuh"
#include "twist_kernels.cuh"

namespace uvs {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    return set_error(UVS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

struct DeviceInfo {
    int device = -1, sms = 0, major = 0, minor = 0, clock_khz = 0;
    size_t smem_optin = 0;
};

static int device_info(DeviceInfo& d) {
    static thread_local DeviceInfo cache;
    int dev = 0;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (cache.device != dev) {
        cudaDeviceProp p;
        if (int rc = check_cuda(cudaGetDeviceProperties(&p, dev), "cudaGetDeviceProperties")) return rc;
        cache.device = dev;
        cache.sms = p.multiProcessorCount;
        cache.major = p.major;
        cache.minor = p.minor;
        cache.smem_optin = p.sharedMemPerBlockOptin;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        cache.clock_khz = khz;
    }
    d = cache;
    return 0;
}

static bool valid_nf(int nf) { return nf == 1 || nf == 3 || nf == 4; }
static bool valid_method(int m) { return m >= UVS_METHOD_KF && m <= UVS_METHOD_GMCKF; }

// kernel family for a request: blockdiag needs identity/NULL Q and R and block P
static int resolve_mode(const uvs_filter_params& fp, bool p_is_blocks_or_identity) {
    if (fp.mode == UVS_MODE_DENSE) return UVS_MODE_DENSE;
    const bool can = (fp.Q == nullptr) && (fp.R == nullptr) && p_is_blocks_or_identity;
    if (fp.mode == UVS_MODE_BLOCKDIAG) return can ? UVS_MODE_BLOCKDIAG : -1;
    return can ? UVS_MODE_BLOCKDIAG : UVS_MODE_DENSE;
}

}  // namespace uvs

#ifdef UVS_PHASE_CLOCKS
static unsigned long long* g_phase_clk = nullptr;      // device array, profiling build only
static unsigned long long* phase_clk_buffer() {
    if (!g_phase_clk) { cudaMalloc((void**)&g_phase_clk, sizeof(unsigned long long) * UVS_NPHASE_BUF); cudaMemset(g_phase_clk, 0, sizeof(unsigned long long) * UVS_NPHASE_BUF); }
    return g_phase_clk;
}
#endif

using namespace uvs;

#pragma GCC visibility push(default)
extern "C" {

int uvs_version(void) { return UVS_ABI_VERSION; }

int uvs_last_error(char* buf, size_t n) {
    if (!buf || n == 0) return UVS_ERR_INVALID;
    strncpy(buf, g_err, n - 1);
    buf[n - 1] = 0;
    return 0;
}

int uvs_device_info(int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz) {
    DeviceInfo d;
    if (int rc = device_info(d)) return rc;
    if (sm_count) *sm_count = d.sms;
    if (cc_major) *cc_major = d.major;
    if (cc_minor) *cc_minor = d.minor;
    if (sm_clock_khz) *sm_clock_khz = d.clock_khz;
    return 0;
}

// range checks shared by the filter entry points: a zero bandwidth or epoch cap would not fail, it would turn every
// kernel weight into NaN (1 / sigma^2 = inf) or skip every correction, and the trial would end FAIL / frozen silently
static int check_filter(const uvs_filter_params& fp, bool needs_sigma) {
    if (!valid_nf(fp.nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", fp.nf);
    const bool mc = fp.method == UVS_METHOD_MCKF || fp.method == UVS_METHOD_IMCCKF || fp.method == UVS_METHOD_GMCKF;
    if (mc && needs_sigma && !(fp.kernel_bw > 0.0) ) return set_error(UVS_ERR_INVALID, "kernel_bw must be > 0 (got %g)", fp.kernel_bw);
    if (fp.method == UVS_METHOD_MCKF) {
        if (fp.fpi_epoch_max < 1) return set_error(UVS_ERR_INVALID, "fpi_epoch_max must be >= 1 (got %d)", fp.fpi_epoch_max);
        if (!(fp.fpi_threshold >= 0.0)) return set_error(UVS_ERR_INVALID, "fpi_threshold must be >= 0 (got %g)", fp.fpi_threshold);
        if (!fp.joseph) return set_error(UVS_ERR_UNSUPPORTED, "MCKF requires the Joseph update (tests/mckf_*.py use it too)");
    }
    if (fp.mode < UVS_MODE_AUTO || fp.mode > UVS_MODE_BLOCKDIAG) return set_error(UVS_ERR_INVALID, "mode %d", fp.mode);
    return 0;
}

static int check_loop_cfg(const uvs_closed_loop_cfg& c) {
    if (c.T < 1) return set_error(UVS_ERR_INVALID, "T=%d", c.T);
    if (!(c.dt > 0.0)) return set_error(UVS_ERR_INVALID, "dt must be > 0 (got %g)", c.dt);
    if (c.sweep_len < 0 || c.sweep_len > UVS_MAX_SWEEP) return set_error(UVS_ERR_INVALID, "sweep_len");
    if (c.sweep_param < 0 || c.sweep_param > 2) return set_error(UVS_ERR_INVALID, "sweep_param must be 0, 1 or 2 (got %d)", c.sweep_param);
    if (c.sweep_len > 0 && c.sweep_epoch < 1) return set_error(UVS_ERR_INVALID, "sweep_epoch must be >= 1 (got %d)", c.sweep_epoch);
    if (c.trial_stride < 0) return set_error(UVS_ERR_INVALID, "trial_stride must be >= 0 (got %d)", c.trial_stride);
    if (!(c.disc_radius >= 0.0)) return set_error(UVS_ERR_INVALID, "disc_radius must be >= 0 (got %g)", c.disc_radius);
    if (c.servo_substeps < 0 || c.servo_substeps > 1000) return set_error(UVS_ERR_INVALID, "servo_substeps must be in [0, 1000] (got %d)", c.servo_substeps);
    if (c.servo_substeps > 0 && !(c.servo_kp > 0.0 && c.servo_kp <= 1.0)) return set_error(UVS_ERR_INVALID, "servo_kp must be in (0, 1] (got %g)", c.servo_kp);
    if (!(c.v_max >= 0.0) || !(c.a_max >= 0.0) || !(c.servo_gain >= 0.0))
        return set_error(UVS_ERR_INVALID, "v_max / a_max / servo_gain must be >= 0");
    return 0;
}

static int run_replay(ReplayArgs& a, void* stream) {
    const uvs_filter_params& fp = a.fp;
    if (!valid_nf(fp.nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", fp.nf);
    if (!valid_method(fp.method)) return set_error(UVS_ERR_INVALID, "method %d is not a filter method", fp.method);
    if (a.B < 0 || a.T < 1) return set_error(UVS_ERR_INVALID, "B=%lld T=%d", a.B, a.T);
    if (a.p_layout != UVS_P_DENSE && a.p_layout != UVS_P_BLOCKS) return set_error(UVS_ERR_INVALID, "p_layout");
    // single step: the bandwidth is the `sigma` argument; replay: kernel_bw (+ annealing)
    if (int rc = check_filter(fp, !a.direct)) return rc;
    const bool mc = fp.method != UVS_METHOD_KF;
    if (a.direct && mc && !(a.sigma > 0.0)) return set_error(UVS_ERR_INVALID, "sigma must be > 0 (got %g)", a.sigma);
    if (a.B == 0) return 0;
    DeviceInfo d;
    if (int rc = device_info(d)) return rc;
    const int m = 2 * fp.nf;
    // the block family needs P given as blocks (or identity) and does not expose K
    const int mode = resolve_mode(fp, (a.p_layout == UVS_P_BLOCKS || a.P == nullptr) && a.K_out == nullptr);
    if (mode < 0) return set_error(UVS_ERR_INVALID, "UVS_MODE_BLOCKDIAG needs Q = R = NULL, a block P and no K_out");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == UVS_MODE_BLOCKDIAG) return check_cuda(launch_blockdiag_replay(m, a, d.sms, st), "blockdiag replay");
    cudaError_t e;
    switch (m) {
    case 2: e = launch_dense_replay<2>(a, d.sms, st); break;
    case 6: e = launch_dense_replay<6>(a, d.sms, st); break;
    default: e = launch_dense_replay<8>(a, d.sms, st); break;
    }
    return check_cuda(e, "dense replay launch");
}

int uvs_filter_step(const uvs_filter_params* p, int64_t B, double* X, double* P, int p_layout,
                    const double* h, const double* Z, double sigma,
                    double* K_out, double* kappa_out, int32_t* flags_out, void* stream) {
    if (!p || (B > 0 && (!X || !P || !h || !Z))) return set_error(UVS_ERR_INVALID, "uvs_filter_step: NULL argument");
    ReplayArgs a{};
    a.fp = *p; a.B = B; a.T = 1; a.k_max = 1; a.direct = 1; a.p_layout = p_layout; a.sigma = sigma;
    a.h = h; a.Z = Z; a.X = X; a.P = P; a.K_out = K_out; a.kappa_out = kappa_out; a.flags_out = flags_out;
    return run_replay(a, stream);
}

int uvs_replay(const uvs_filter_params* p, int64_t B, int32_t T, int32_t k_max,
               const double* f, const double* dq, const double* f_init,
               double* X, double* P, int p_layout,
               double* X_out, double* P_out, double* kappa_out, int32_t* flags_out, void* stream) {
    if (!p || (B > 0 && (!f || !dq || !f_init || !X))) return set_error(UVS_ERR_INVALID, "uvs_replay: NULL argument");
    ReplayArgs a{};
    a.fp = *p; a.B = B; a.T = T; a.k_max = k_max > 0 ? k_max : T; a.direct = 0; a.p_layout = p_layout;
    a.f = f; a.dq = dq; a.f_init = f_init; a.X = X; a.P = P;
    a.X_out = X_out; a.P_out = P_out; a.kappa_out = kappa_out; a.flags_out = flags_out;
    return run_replay(a, stream);
}

int uvs_ibvs(int32_t nf, int64_t B, const double* J, const double* e, const double* kappa, double gain,
             double* dq_out, int32_t* flags_out, void* stream) {
    if (!valid_nf(nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", nf);
    if (B < 0 || (B > 0 && (!J || !e || !dq_out))) return set_error(UVS_ERR_INVALID, "uvs_ibvs: bad argument");
    if (B == 0) return 0;
    return check_cuda(launch_ibvs(2 * nf, B, J, e, kappa, gain, dq_out, flags_out, (cudaStream_t)stream), "ibvs");
}

int uvs_ur10_fk(int64_t B, const double* q, double* T06_out, double* Jgeom_out, void* stream) {
    if (B < 0 || (B > 0 && !q)) return set_error(UVS_ERR_INVALID, "uvs_ur10_fk: bad argument");
    if (B == 0) return 0;
    return check_cuda(launch_fk(B, q, T06_out, Jgeom_out, (cudaStream_t)stream), "ur10_fk");
}

int uvs_project(int32_t nf, int64_t B, const double* q, const double* targets_host,
                double* f_out, double* depth_out, void* stream) {
    if (!valid_nf(nf) || !targets_host || B < 0 || (B > 0 && !q)) return set_error(UVS_ERR_INVALID, "uvs_project: bad argument");
    if (B == 0) return 0;
    return check_cuda(launch_project(nf, B, q, targets_host, f_out, depth_out, nullptr, (cudaStream_t)stream),
                      "project");
}

int uvs_initial_guess(int32_t nf, int64_t B, const double* q, const double* targets_host,
                      double* X0_out, double* f_out, void* stream) {
    if (!valid_nf(nf) || !targets_host || B < 0 || (B > 0 && (!q || !X0_out)))
        return set_error(UVS_ERR_INVALID, "uvs_initial_guess: bad argument");
    if (B == 0) return 0;
    return check_cuda(launch_project(nf, B, q, targets_host, f_out, nullptr, X0_out, (cudaStream_t)stream),
                      "initial_guess");
}

static int check_noise(const uvs_noise_params& np) {
    if (np.type < UVS_NOISE_NONE || np.type > UVS_NOISE_UNIFORM) return set_error(UVS_ERR_INVALID, "noise type %d", np.type);
    if (np.type == UVS_NOISE_ALPHA_STABLE && !(np.alpha > 0.0 && np.alpha <= 2.0))
        return set_error(UVS_ERR_INVALID, "alpha-stable needs 0 < alpha <= 2 (got %g)", np.alpha);
    if (np.hold_cnt < 0) return set_error(UVS_ERR_INVALID, "hold_cnt < 0");
    return 0;
}

int uvs_noise_fill(const uvs_noise_params* np, uint64_t trial0, int64_t B, int32_t T, int32_t m,
                   double* out, void* stream) {
    if (!np || B < 0 || T < 0 || m < 2 || (m & 1) || (B > 0 && T > 0 && !out)) return set_error(UVS_ERR_INVALID, "uvs_noise_fill: bad argument");
    if (int rc = check_noise(*np)) return rc;
    if (B == 0 || T == 0) return 0;
    return check_cuda(launch_noise_fill(*np, trial0, B, T, m, out, (cudaStream_t)stream), "noise_fill");
}

int uvs_start_poses(const double* q_start_host, uint64_t pose_seed, int32_t change_q_start,
                    uint64_t trial0, int64_t B, double* q_out, void* stream) {
    if (!q_start_host || B < 0 || (B > 0 && !q_out)) return set_error(UVS_ERR_INVALID, "uvs_start_poses: bad argument");
    if (B == 0) return 0;
    return check_cuda(launch_start_poses(q_start_host, pose_seed, change_q_start, trial0, B, q_out,
                                         (cudaStream_t)stream), "start_poses");
}

int uvs_closed_loop(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B,
                    const uvs_closed_loop_io* io, void* stream) {
    if (!cfg || !io || (B > 0 && (!io->stats || !io->status))) return set_error(UVS_ERR_INVALID, "uvs_closed_loop: NULL argument");
    const uvs_filter_params& fp = cfg->filt;
    if (!valid_nf(fp.nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", fp.nf);
    if (!valid_method(fp.method)) return set_error(UVS_ERR_INVALID, "method %d is not a filter method", fp.method);
    if (B < 0) return set_error(UVS_ERR_INVALID, "B=%lld", (long long)B);
    if (int rc = check_loop_cfg(*cfg)) return rc;
    if (int rc = check_filter(fp, true)) return rc;
    if (io->p_layout != UVS_P_DENSE && io->p_layout != UVS_P_BLOCKS) return set_error(UVS_ERR_INVALID, "p_layout");
    if (int rc = check_noise(cfg->noise)) return rc;
    if (B == 0) return 0;
    DeviceInfo d;
    if (int rc = device_info(d)) return rc;
    const int m = 2 * fp.nf;
    const int mode = resolve_mode(fp, cfg->P0 == nullptr);
    if (mode < 0) return set_error(UVS_ERR_INVALID, "UVS_MODE_BLOCKDIAG needs Q = R = P0 = NULL");
    ClosedLoopArgs a{};
    a.cfg = *cfg; a.io = *io; a.trial0 = trial0; a.B = B;
    a.stride = cfg->trial_stride > 1 ? (unsigned long long)cfg->trial_stride : 1ull;
#ifdef UVS_PHASE_CLOCKS
    a.phase_clk = phase_clk_buffer();
#endif
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == UVS_MODE_BLOCKDIAG) return check_cuda(launch_blockdiag_closed_loop(m, a, d.sms, st), "blockdiag closed loop");
    cudaError_t e;
    switch (m) {
    case 2: e = launch_dense_closed_loop<2>(a, d.sms, st); break;
    case 6: e = launch_dense_closed_loop<6>(a, d.sms, st); break;
    default: e = launch_dense_closed_loop<8>(a, d.sms, st); break;
    }
    return check_cuda(e, "dense closed-loop launch");
}

int uvs_twist_loop(const uvs_closed_loop_cfg* cfg, const uvs_twist_cfg* tw, uint64_t trial0, int64_t B,
                   const uvs_closed_loop_io* io, void* stream) {
    if (!cfg || !tw || !io || (B > 0 && (!io->stats || !io->status))) return set_error(UVS_ERR_INVALID, "uvs_twist_loop: NULL argument");
    const uvs_filter_params& fp = cfg->filt;
    if (!valid_nf(fp.nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", fp.nf);
    if (fp.method != UVS_METHOD_ANALYTICAL && !valid_method(fp.method)) return set_error(UVS_ERR_INVALID, "method %d", fp.method);
    if (B < 0) return set_error(UVS_ERR_INVALID, "B=%lld", (long long)B);
    if (int rc = check_loop_cfg(*cfg)) return rc;
    if (int rc = check_filter(fp, true)) return rc;
    if (io->p_layout != UVS_P_DENSE && io->p_layout != UVS_P_BLOCKS) return set_error(UVS_ERR_INVALID, "p_layout");
    if (tw->x0_mode < UVS_X0_ZEROS || tw->x0_mode > UVS_X0_RANDOM) return set_error(UVS_ERR_INVALID, "x0_mode %d", tw->x0_mode);
    if (tw->depth_mode != UVS_DEPTH_DISTANCE && tw->depth_mode != UVS_DEPTH_HEIGHT) return set_error(UVS_ERR_INVALID, "depth_mode %d", tw->depth_mode);
    if (tw->dp_zero_mask < 0 || tw->dp_zero_mask > 63) return set_error(UVS_ERR_INVALID, "dp_zero_mask %d", tw->dp_zero_mask);
    if ((fp.method == UVS_METHOD_ANALYTICAL && !(tw->focal > 0.0)) ||
        (tw->x0_mode == UVS_X0_IMAGE_JACOBIAN && !(tw->focal_x0 > 0.0)))
        return set_error(UVS_ERR_INVALID, "focal lengths must be positive");
    if (fp.mode == UVS_MODE_BLOCKDIAG) return set_error(UVS_ERR_UNSUPPORTED, "the twist loop runs in the dense family only");
    if (int rc = check_noise(cfg->noise)) return rc;
    if (B == 0) return 0;
    DeviceInfo d;
    if (int rc = device_info(d)) return rc;
    TwistArgs a{};
    a.cfg = *cfg; a.tw = *tw; a.io = *io; a.trial0 = trial0; a.B = B;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    switch (2 * fp.nf) {
    case 2: e = launch_twist_loop<2>(a, d.sms, st); break;
    case 6: e = launch_twist_loop<6>(a, d.sms, st); break;
    default: e = launch_twist_loop<8>(a, d.sms, st); break;
    }
    return check_cuda(e, "twist loop launch");
}

int uvs_closed_loop_host(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B,
                         double* stats_host, int32_t* status_host, int32_t* frames_host) {
    if (!cfg || B < 0 || (B > 0 && (!stats_host || !status_host))) return set_error(UVS_ERR_INVALID, "uvs_closed_loop_host: bad argument");
    if (B == 0) return 0;
    // results land in a per-host-thread device scratch block that only grows (a stream-ordered cudaMallocAsync /
    // cudaFreeAsync pair per call was measured to stall for up to 1.8 s when the pool hands memory back to the driver
    // at the synchronisation point), are copied straight into the caller's buffers (a DMA when they are page-locked)
    // and the call ends with one stream synchronisation
    const size_t n_stats = sizeof(double) * 3 * (size_t)B, n_i32 = sizeof(int32_t) * (size_t)B;
    cudaStream_t st = cudaStreamPerThread;
    struct Scratch {
        unsigned char* p = nullptr; size_t cap = 0; int dev = -1;
        ~Scratch() {                                            // the owning host thread exits: give the block back
            if (!p) return;
            int cur = 0;
            if (cudaGetDevice(&cur) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) { cudaFree(p); cudaSetDevice(cur); }
        }
    };
    constexpr int kMaxDevices = 64;
    static thread_local Scratch per_device[kMaxDevices];       // one block per (host thread, device)
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    if (dev < 0 || dev >= kMaxDevices) return set_error(UVS_ERR_UNSUPPORTED, "device ordinal %d", dev);
    Scratch& sc = per_device[dev];
    if (sc.cap < n_stats + 2 * n_i32) {
        if (sc.p) cudaFree(sc.p);
        sc.p = nullptr; sc.cap = 0; sc.dev = dev;
        rc = check_cuda(cudaMalloc((void**)&sc.p, n_stats + 2 * n_i32), "cudaMalloc result scratch");
        if (rc) { sc.p = nullptr; return rc; }
        sc.cap = n_stats + 2 * n_i32;
    }
    unsigned char* scratch = sc.p;
    double* d_stats = reinterpret_cast<double*>(scratch);
    int32_t* d_status = reinterpret_cast<int32_t*>(scratch + n_stats);
    int32_t* d_frames = reinterpret_cast<int32_t*>(scratch + n_stats + n_i32);
    {
        uvs_closed_loop_io io{};
        io.stats = d_stats; io.status = d_status; io.frames = d_frames;
        rc = uvs_closed_loop(cfg, trial0, B, &io, st);
    }
    if (!rc) rc = check_cuda(cudaMemcpyAsync(stats_host, d_stats, n_stats, cudaMemcpyDeviceToHost, st), "D2H stats");
    if (!rc) rc = check_cuda(cudaMemcpyAsync(status_host, d_status, n_i32, cudaMemcpyDeviceToHost, st), "D2H status");
    if (!rc && frames_host)
        rc = check_cuda(cudaMemcpyAsync(frames_host, d_frames, n_i32, cudaMemcpyDeviceToHost, st), "D2H frames");
    const int rc_sync = check_cuda(cudaStreamSynchronize(st), "uvs_closed_loop_host");
    return rc ? rc : rc_sync;
}

int uvs_closed_loop_stream(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B, int64_t chunk,
                           uvs_log_sink sink, void* user, double* d2h_gbs, double* wall_s) {
    if (!cfg || B < 0 || chunk < 1) return set_error(UVS_ERR_INVALID, "uvs_closed_loop_stream: bad argument");
    if (!valid_nf(cfg->filt.nf)) return set_error(UVS_ERR_INVALID, "nf must be 1, 3 or 4 (got %d)", cfg->filt.nf);
    if (cfg->T < 1) return set_error(UVS_ERR_INVALID, "T=%d", cfg->T);
    if (d2h_gbs) *d2h_gbs = 0.0;
    if (wall_s) *wall_s = 0.0;
    if (B == 0) return 0;
    if (chunk > B) chunk = B;
    const int m = 2 * cfg->filt.nf, logw = 6 + 2 * m;
    const size_t n_log = sizeof(double) * (size_t)chunk * cfg->T * logw, n_stats = sizeof(double) * 3 * (size_t)chunk,
                 n_i32 = sizeof(int32_t) * (size_t)chunk;
    const size_t bytes = n_log + n_stats + 2 * n_i32;
    struct Slot {
        unsigned char *dev = nullptr, *host = nullptr;
        cudaEvent_t computed = nullptr, copy_begin = nullptr, copied = nullptr;
        long long first = 0, n = 0;
    } slot[2];
    cudaStream_t compute = nullptr, copy = nullptr;
    int rc = 0;
    auto fail = [&](cudaError_t e, const char* what) { if (!rc && e != cudaSuccess) rc = check_cuda(e, what); return rc != 0; };
    fail(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking), "stream");
    fail(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking), "stream");
    for (int i = 0; i < 2 && !rc; ++i) {
        fail(cudaMalloc((void**)&slot[i].dev, bytes), "cudaMalloc log slot");
        fail(cudaHostAlloc((void**)&slot[i].host, bytes, cudaHostAllocDefault), "cudaHostAlloc log slot");
        fail(cudaEventCreate(&slot[i].computed), "event");
        fail(cudaEventCreate(&slot[i].copy_begin), "event");
        fail(cudaEventCreate(&slot[i].copied), "event");
    }
    auto ptrs = [&](unsigned char* base, double*& log, double*& stats, int32_t*& status, int32_t*& frames) {
        log = reinterpret_cast<double*>(base);
        stats = reinterpret_cast<double*>(base + n_log);
        status = reinterpret_cast<int32_t*>(base + n_log + n_stats);
        frames = reinterpret_cast<int32_t*>(base + n_log + n_stats + n_i32);
    };
    double copy_ms = 0.0, copied_bytes = 0.0;
    // drain slot s: wait for its copy, account the copy-engine time, hand the chunk to the sink
    auto drain = [&](Slot& s) {
        if (s.n == 0 || rc) return;
        if (fail(cudaEventSynchronize(s.copied), "log D2H")) return;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.copy_begin, s.copied);
        copy_ms += ms;
        double *log, *stats; int32_t *status, *frames;
        ptrs(s.host, log, stats, status, frames);
        copied_bytes += (double)(sizeof(double) * (size_t)s.n * cfg->T * logw + sizeof(double) * 3 * s.n + 2 * sizeof(int32_t) * s.n);
        if (sink && sink(user, (uint64_t)s.first, s.n, log, stats, status, frames) != 0)
            rc = set_error(UVS_ERR_INVALID, "uvs_closed_loop_stream: the sink asked to stop at trial %lld", s.first);
        s.n = 0;
    };
    const auto t_begin = std::chrono::steady_clock::now();        // after the allocations: the steady-state pipeline
    long long done = 0;
    for (int it = 0; done < B && !rc; ++it) {
        Slot& s = slot[it & 1];
        drain(s);                                       // its previous chunk (two launches ago) must have left the buffers
        if (rc) break;
        const long long n = (B - done < chunk) ? (B - done) : chunk;
        double *log, *stats; int32_t *status, *frames;
        ptrs(s.dev, log, stats, status, frames);
        uvs_closed_loop_io io{};
        io.stats = stats; io.status = status; io.frames = frames; io.log = log;
        if ((rc = uvs_closed_loop(cfg, trial0 + (uint64_t)done, n, &io, compute)) != 0) break;
        fail(cudaEventRecord(s.computed, compute), "event record");
        fail(cudaStreamWaitEvent(copy, s.computed, 0), "stream wait");
        fail(cudaEventRecord(s.copy_begin, copy), "event record");
        // one contiguous copy: the chunk's log, stats, status and frames are laid out back to back (a short last chunk
        // copies its own log prefix and the three small arrays separately)
        if (n == chunk) fail(cudaMemcpyAsync(s.host, s.dev, bytes, cudaMemcpyDeviceToHost, copy), "log D2H");
        else {
            fail(cudaMemcpyAsync(s.host, s.dev, sizeof(double) * (size_t)n * cfg->T * logw, cudaMemcpyDeviceToHost, copy), "log D2H");
            fail(cudaMemcpyAsync(s.host + n_log, s.dev + n_log, n_stats + 2 * n_i32, cudaMemcpyDeviceToHost, copy), "log D2H");
        }
        fail(cudaEventRecord(s.copied, copy), "event record");
        s.first = (long long)trial0 + done;
        s.n = n;
        done += n;
        drain(slot[(it + 1) & 1]);                      // the other slot's chunk: sink it while this one computes
    }
    drain(slot[0]);
    drain(slot[1]);
    if (!rc) {
        fail(cudaStreamSynchronize(copy), "uvs_closed_loop_stream");
        fail(cudaStreamSynchronize(compute), "uvs_closed_loop_stream");
        if (wall_s) *wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
        if (d2h_gbs && copy_ms > 0.0) *d2h_gbs = copied_bytes / (copy_ms * 1e-3) / 1e9;
    } else {
        cudaStreamSynchronize(copy);
        cudaStreamSynchronize(compute);
    }
    for (int i = 0; i < 2; ++i) {
        if (slot[i].dev) cudaFree(slot[i].dev);
        if (slot[i].host) cudaFreeHost(slot[i].host);
        if (slot[i].computed) cudaEventDestroy(slot[i].computed);
        if (slot[i].copy_begin) cudaEventDestroy(slot[i].copy_begin);
        if (slot[i].copied) cudaEventDestroy(slot[i].copied);
    }
    if (compute) cudaStreamDestroy(compute);
    if (copy) cudaStreamDestroy(copy);
    return rc;
}

int uvs_reduce_stats(int64_t B, uint64_t trial0, int32_t trial_stride, int32_t sweep_epoch, int32_t n_levels,
                     const double* stats, const int32_t* status, double* sums, double* n_fail, void* stream) {
    // n_levels * 80 B of shared memory per CTA: 512 levels stay under the 48 KB a kernel gets without opting in
    if (!sums || !n_fail || B < 0 || n_levels < 1 || n_levels > 512 || (B > 0 && (!stats || !status)))
        return set_error(UVS_ERR_INVALID, "uvs_reduce_stats: bad argument (1 <= n_levels <= 512)");
    if (B == 0) return 0;
    if (trial_stride < 0) return set_error(UVS_ERR_INVALID, "uvs_reduce_stats: trial_stride < 0");
    return check_cuda(launch_reduce_stats(B, trial0, trial_stride, sweep_epoch, n_levels, stats, status, sums, n_fail,
                                          (cudaStream_t)stream), "reduce_stats");
}

#ifdef UVS_PHASE_CLOCKS
// profiling build only: read (and reset) the per-phase clock totals of the dense closed-loop kernel
int uvs_debug_phase_clocks(unsigned long long* out_host, int reset) {
    cudaDeviceSynchronize();
    if (out_host) cudaMemcpy(out_host, phase_clk_buffer(), sizeof(unsigned long long) * UVS_NPHASE_BUF, cudaMemcpyDeviceToHost);
    if (reset) cudaMemset(phase_clk_buffer(), 0, sizeof(unsigned long long) * UVS_NPHASE_BUF);
    return 0;
}
#endif

int uvs_fp64_peak(int32_t iters, double* tflops_host, double* ms_host) {
    if (!tflops_host || iters < 1) return set_error(UVS_ERR_INVALID, "uvs_fp64_peak: bad argument");
    DeviceInfo d;
    if (int rc = device_info(d)) return rc;
    double ms = 0.0;
    int rc = check_cuda(run_fp64_peak(iters, d.sms, tflops_host, &ms), "fp64 peak probe");
    if (ms_host) *ms_host = ms;
    return rc;
}

}  // extern "C"
#pragma GCC visibility pop
