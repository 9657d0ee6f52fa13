/*
 * uvs_b200.h — C ABI of the B200-native Monte Carlo engine for the
 * uncalibrated-visual-servoing hot path (per-frame Jacobian estimators + IBVS).
 *
 * The reference (AI-SPARC/uncalibrated-visual-servoing) has no FFI: its
 * estimators are inline blocks of Experiment.run() (experiment.py:48-359).
 * Each entry point below names the reference lines it replaces; a reference
 * maintainer binds them with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success or a negative UVS_ERR_* code; the
 *     message of the last failure (per host thread) is read with uvs_last_error();
 *   - no ownership transfer: the caller allocates every buffer;
 *   - pointers are DEVICE pointers unless the name ends in _host or the comment
 *     says "host"; structs are passed by host pointer and read synchronously;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 *     are asynchronous on it unless stated otherwise;
 *   - an empty batch (B == 0) is a successful no-op after the scalar arguments have been
 *     validated; its batch buffers may be NULL (an empty tensor has no storage);
 *   - all arithmetic is IEEE FP64; matrices are row-major and contiguous;
 *   - nf = number of image features (1, 3 or 4); m = 2*nf measurements;
 *     N = 6*m states (the row-major vectorised m x 6 feature Jacobian).
 */
#ifndef UVS_B200_H
#define UVS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UVS_ABI_VERSION 4

/* experiment.py:6-11 (Method).  RMCKF is the paper's name for GMCKF. */
enum { UVS_METHOD_ANALYTICAL = 1, UVS_METHOD_KF = 2, UVS_METHOD_MCKF = 3,
       UVS_METHOD_IMCCKF = 4, UVS_METHOD_GMCKF = 5, UVS_METHOD_RMCKF = 5 };

/* noise.py:7-12 (NoiseType); 0 = no noise, 6 = caller-supplied stream */
enum { UVS_NOISE_NONE = 0, UVS_NOISE_WHITE = 1, UVS_NOISE_GAUSSIAN_MIXTURE = 2,
       UVS_NOISE_GAUSSIAN_BIMODAL = 3, UVS_NOISE_ALPHA_STABLE = 4, UVS_NOISE_UNIFORM = 5 };

/* experiment.py:13-15 (ExperimentStatus) */
enum { UVS_STATUS_SUCCESS = 0, UVS_STATUS_FAIL = 1 };

/* covariance representation / kernel family */
enum { UVS_MODE_AUTO = 0,       /* blockdiag when P0,Q,R allow it, else dense            */
       UVS_MODE_DENSE = 1,      /* general SPD P (N x N), Q, R                           */
       UVS_MODE_BLOCKDIAG = 2 };/* P = m diagonal 6x6 blocks, R diagonal (exact for the  */
                                /* reference's hard-coded P0=Q=I, R=I; experiment.py:73-76) */

/* layout of P buffers at the ABI */
enum { UVS_P_DENSE = 0,         /* [B, N, N]                                              */
       UVS_P_BLOCKS = 1 };      /* [B, m, 6, 6] diagonal blocks only                      */

/* per-frame flags (bit-or); epoch count of the MCKF fixed point in bits 8.. */
#define UVS_FLAG_SKIP_SINGULAR   1   /* MCKF: a measurement weight underflowed to 0 -> correction skipped (experiment.py:233-236) */
#define UVS_FLAG_SKIP_MAXEPOCH   2   /* MCKF: fpi_epoch_max reached -> correction skipped (experiment.py:246-248) */
#define UVS_FLAG_CX_SINGULAR     4   /* MCKF: a state weight underflowed (the reference raises out of run() at experiment.py:224);
                                        closed loops: the trial ends UVS_STATUS_FAIL on that frame */
#define UVS_FLAG_NOT_SPD         8   /* Cholesky of P or S met a non-positive pivot -> correction skipped, the trial goes on (the
                                        reference inverts S without a Cholesky; P0, Q, R must be SPD) */
#define UVS_FLAG_RANK_DEFICIENT 16   /* IBVS: cond(J) >~ 1e5, pseudo-inverse by one-sided Jacobi on J with numpy's 1e-15 sigma_max cutoff */
#define UVS_FLAG_NONFINITE      32   /* trial failed: non-finite state / command (experiment.py:312-316) */
#define UVS_FLAG_OUT_OF_VIEW    64   /* trial failed: a feature left the field of view (utils.py:73-78 -> NaN centroid) */
#define UVS_FLAG_EPOCH_SHIFT     8

#define UVS_ERR_INVALID   (-1)  /* bad argument */
#define UVS_ERR_CUDA      (-2)  /* CUDA runtime error (message has cudaGetErrorString) */
#define UVS_ERR_UNSUPPORTED (-3)

#define UVS_MAX_SWEEP 64

/* estimator parameters: config.json "estimator" block + experiment.py:70-76 */
typedef struct uvs_filter_params {
    int32_t method;          /* UVS_METHOD_*                                              */
    int32_t nf;              /* 1, 3, 4                                                   */
    int32_t mode;            /* UVS_MODE_*                                                */
    int32_t joseph;          /* 1: P=(I-KH)P(I-KH)^T+KRK^T (experiment.py:297); 0: P=(I-KH)P (tests/kalman_1_feature.py:126) */
    int32_t mckf_skip;       /* 1: singular / max-epoch skip (experiment.py:233-248); 0: tests/mckf_* behaviour */
    int32_t fpi_epoch_max;   /* experiment.py:39                                          */
    int32_t annealing;       /* experiment.py:40,196-199: sigma = kernel_bw + 100(1-k/k_max) */
    int32_t reserved;
    double kernel_bw;        /* experiment.py:37                                          */
    double fpi_threshold;    /* experiment.py:38                                          */
    const double* Q;         /* device [N,N] process covariance, NULL = identity (experiment.py:75) */
    const double* R;         /* device [m,m] measurement covariance, NULL = identity (experiment.py:76) */
} uvs_filter_params;

/* noise parameters: config.json "noise" block + noise.py:31-79 */
typedef struct uvs_noise_params {
    int32_t type;            /* UVS_NOISE_*                                               */
    int32_t hold_cnt;        /* noise_hold_cnt_max_predefined (0 = no hold), noise.py:72-75 */
    double std, mean, rho;   /* Gaussian types, noise.py:46-51                            */
    double alpha, beta, gamma, delta;   /* alpha-stable, noise.py:60-64                   */
    uint64_t seed;           /* Philox key                                                */
} uvs_noise_params;

/* one closed-loop Monte Carlo batch: experiment.py:48-359 + main.py:104-139 */
typedef struct uvs_closed_loop_cfg {
    uvs_filter_params filt;
    uvs_noise_params noise;
    int32_t T;               /* frames per trial = int(t_max/dt), frame k logged at t=(k+1)dt */
    int32_t initial_guess;   /* experiment.py:86-117                                      */
    int32_t change_q_start;  /* main.py:129-134                                           */
    int32_t fov_check;       /* 1: a feature outside [0,256) or depth<=0.01 fails the trial */
    int32_t sweep_len;       /* number of noise levels (main.py:104-106), 0 = no sweep    */
    int32_t sweep_epoch;     /* trials per noise level (config "epoch"); level = trial / sweep_epoch */
    int32_t sweep_param;     /* 0 none, 1 = rho, 2 = alpha (main.py:123-126)              */
    int32_t trial_stride;    /* entry b of the batch is global trial trial0 + b * trial_stride (0 or 1: contiguous).  Rank r
                                of G takes trial0 = r, trial_stride = G to get the same mix of noise levels as every other
                                rank: a frame's cost depends on the level (alpha = 1 and 2 use cheap samplers), so contiguous
                                shards of a sweep finish at different times (measured: 2.7 % at G = 2)               */
    double dt;               /* config "dt"                                               */
    double ibvs_gain;        /* config "ibvs_gain"                                        */
    double q_start[6];
    double desired_f[8];
    double targets[12];      /* nf world points (x,y,z), feature order red, green, blue, pink */
    uint64_t pose_seed;      /* Philox key of the start-pose draws (config experiments.seed) */
    double sweep[UVS_MAX_SWEEP];
    const double* P0;        /* device [N,N] or NULL = identity (experiment.py:73)         */
    /* joint servo limits of the scenes (position-controlled joints, Ruc v_max = 1.0472 rad/s,
     * a_max = 6.283 rad/s^2); 0 = unlimited = the kinematic plant q <- q + dq dt of experiment.py:320,335.
     * servo_gain: fraction of the commanded step a position-controlled joint covers within one frame (first-order
     * servo lag; the scenes' joints run a P controller, gain 0.1, at a 5 ms dynamics step); 0 or 1 = exact tracking.
     * Per frame: v = servo_gain dq ; v = clamp(v, v_prev -+ a_max dt) ; v = clamp(v, -+v_max) ; q += v dt ; v_prev = v.
     * The filter regressor stays the COMMANDED dq, as in the reference (experiment.py:188). */
    double v_max;
    double a_max;
    double servo_gain;
    /* Scene plant (constants decoded from scenarios/*.ttt; every field 0 = off = the ideal
     * kinematic point-feature plant).  Launches that use any of them run the general kernel instantiations.
     * sensor_offset: origin of the vision sensor in DH frame 6 (the scenes mount it ~(0.004, 0.007, 0.014) m off).
     * disc_radius:   the targets are discs of this radius lying in the plane z = const (scenes: 0.0375 m); a feature is then
     *                the CENTROID OF THE DISC'S IMAGE REGION, which is what utils.py:73-78 measures (255 sum(X mask) / sum(mask)),
     *                evaluated on the 16-gon through the projected rim; with fov_check the region is clipped to the image
     *                [-0.5, 255.5]^2 (a partly visible disc gives a biased centroid, as in the reference) and the trial fails
     *                only when NOTHING of a disc is visible (0/0 -> NaN -> pinv raises, experiment.py:312-316).
     * servo_substeps / servo_kp: the joints are position-controlled (P gain 0.1) at the 5 ms dynamics step, 10 substeps per
     *                frame: per substep h = dt / substeps, v = kp (q_target - q) / h, limited by a_max h and v_max, q += v h,
     *                with q_target = q + dq dt (experiment.py:320).  Without limits this equals servo_gain = 1 - (1 - kp)^substeps. */
    double sensor_offset[3];
    double disc_radius;
    double servo_kp;
    int32_t servo_substeps;
    int32_t reserved2;
} uvs_closed_loop_cfg;

/* buffers of a closed-loop batch; any optional pointer may be NULL */
typedef struct uvs_closed_loop_io {
    double*  stats;          /* [B,3]  ISE, IAE, ITAE per trial (results/plot_errorbar.m:39-84) */
    int32_t* status;         /* [B]    UVS_STATUS_*                                        */
    int32_t* frames;         /* [B]    frames completed                                    */
    int32_t* flags;          /* [B]    optional: bit-or of all per-frame flags             */
    double*  log;            /* optional [B,T,6+2m]: q(6), f(m) noisy, noise(m) per frame (experiment.py:323-328) */
    double*  X_out;          /* optional [B,N] final state                                 */
    double*  P_out;          /* optional final covariance, layout p_layout                 */
    const double* noise_in;  /* optional [B,T,m] noise stream replacing the Philox draws   */
    const double* q0_in;     /* optional [B,6] start poses replacing main.py:129-134       */
    const double* X0_in;     /* optional [B,N] initial state replacing experiment.py:86-117 */
    int32_t p_layout;        /* UVS_P_* of P_out                                           */
    int32_t reserved;
    /* Per-frame trace of the closed loop (ABI 4) — the open-loop replay gate of the closed-loop kernels themselves:
     * with dq_in the command stream of a recorded run REPLACES the command the IBVS law computed (as the regressor of
     * the next frame, experiment.py:188, and in the joint integration, experiment.py:320), so frame k of the kernel
     * sees the recorded run's inputs and its X / P / dq / kappa can be compared frame by frame (1e-9) with what the
     * reference computed; the computed command is still what dq_log receives.  All optional. */
    const double* dq_in;     /* [B,T,6] recorded commands; ignored by uvs_twist_loop (its regressor is dp)  */
    double*  X_log;          /* [B,T,N] state after the update of frame k (experiment.py:300)              */
    double*  P_log;          /* [B,T,m,6,6] diagonal 6x6 blocks of the covariance after the update of frame k */
    double*  dq_log;         /* [B,T,6] command computed in frame k (experiment.py:312; twist loop: joint rates) */
    double*  kappa_log;      /* [B,T,m] weights kappa of frame k (experiment.py:306-308)                   */
    int32_t* flags_log;      /* [B,T] per-frame flags | epochs << UVS_FLAG_EPOCH_SHIFT                     */
} uvs_closed_loop_io;

/* The camera-twist pipeline of the reference's tests/ scripts (tests/{ibvs,kalman,mckf}_{1,3,4}_feature(s).py):
 * the filter regresses the feature differences on the commanded camera twist dp, the control law
 * produces dp = -GAIN pinv(J)(f - f*), and the joints follow dq = pinv(J_geom) kron(I2, R06^T) dp
 * (kalman_4_features.py:152-173).  GAIN, TS, T, q_start, desired_f, targets, noise and the filter
 * parameters are the uvs_closed_loop_cfg fields (ibvs_gain, dt, ...); this block adds what only the scripts have. */
enum { UVS_X0_ZEROS = 0,            /* X = 0 (kalman_1_feature.py:44); first f_old = 0               */
       UVS_X0_IMAGE_JACOBIAN = 1,   /* X = analytic interaction matrix at the start pose, first f_old =
                                       the noise-free features (kalman_4_features.py:74-101)         */
       UVS_X0_RANDOM = 2 };         /* X ~ U(0,1)^N from Philox (kalman_3_features.py:50); f_old = 0 */
enum { UVS_DEPTH_DISTANCE = 0,      /* Z_i = |camera - disc_i| (computeZ, ur10_simulation.py:170-195) */
       UVS_DEPTH_HEIGHT = 1 };      /* Z = T06[2][3] (getCameraHeight, ibvs_1_feature.py:58)          */

typedef struct uvs_twist_cfg {
    int32_t x0_mode;         /* UVS_X0_*; io->X0_in replaces the state but not the first f_old     */
    int32_t depth_mode;      /* UVS_DEPTH_* used by the interaction matrix                        */
    int32_t dp_zero_mask;    /* bit i set: dp[i] = 0 after the control law (kalman_3_features.py:139-141) */
    int32_t stale_fk;        /* 1: R06 and J_geom of the START pose are used in every frame — what kalman_4_features.py:165-169
                                and mckf_{1,3,4}_feature(s).py do: they call getCameraRotation() / jacobian() without
                                recalculate_fkine, and UR10Simulation only refreshes its cached transforms on
                                fkine(recalculate=True) (ur10_simulation.py:97-110,141-144); 0: the pose of the frame
                                (kalman_{1,3}_feature(s).py:133/142 and ibvs_*.py refresh it every frame) */
    double focal_x0;         /* focal length of the initial guess (kalman_4_features.py:85: 1024)  */
    double focal;            /* focal length of the UVS_METHOD_ANALYTICAL interaction matrix (ibvs_4_features.py:81) */
    double err_threshold;    /* frames run while ||f - f*||_2 > err_threshold (kalman_4_features.py:111); < 0: all T */
} uvs_twist_cfg;

/* library / error plumbing */
int uvs_version(void);
int uvs_last_error(char* buf, size_t n);
int uvs_device_info(int* sm_count, int* cc_major, int* cc_minor, int* sm_clock_khz);   /* host outputs */

/* One predict+update of B independent filters — experiment.py:166-297
 * (tests/kalman_*: joseph=0; tests/mckf_*: mckf_skip=0).
 * X [B,N] in/out; P in/out (p_layout); h [B,6] regressor (zeros on a first
 * frame, experiment.py:183-188); Z [B,m]; sigma = kernel bandwidth of this frame.
 * Optional outputs: K_out [B,N,m], kappa_out [B,m] (experiment.py:306-308),
 * flags_out [B] (flags | epochs << UVS_FLAG_EPOCH_SHIFT). */
int uvs_filter_step(const uvs_filter_params* p, int64_t B, double* X, double* P, int p_layout,
                    const double* h, const double* Z, double sigma,
                    double* K_out, double* kappa_out, int32_t* flags_out, void* stream);

/* T frames of B filters without leaving the device — the open-loop replay of
 * experiment.py:125-300: frame k uses Z = f[k]-f[k-1] (f[-1] = f_init) and the
 * regressor dq[k-1] (zeros at k=0).  sigma follows the annealing rule with
 * frame index k and k_max.  X [B,N], P in/out.  Optional per-frame outputs:
 * X_out [B,T,N], P_out [B,T,...] (p_layout), kappa_out [B,T,m], flags_out [B,T]. */
int uvs_replay(const uvs_filter_params* p, int64_t B, int32_t T, int32_t k_max,
               const double* f, const double* dq, const double* f_init,
               double* X, double* P, int p_layout,
               double* X_out, double* P_out, double* kappa_out, int32_t* flags_out, void* stream);

/* IBVS law dq = -gain * pinv(J) (kappa o e) — experiment.py:300-312.
 * J [B,m,6], e [B,m], kappa [B,m] or NULL (= ones); dq_out [B,6]; flags_out [B] or NULL. */
int uvs_ibvs(int32_t nf, int64_t B, const double* J, const double* e, const double* kappa, double gain,
             double* dq_out, int32_t* flags_out, void* stream);

/* UR10 DH forward kinematics + geometric Jacobian — ur10_simulation.py:97-139.
 * q [B,6]; T06_out [B,16] row-major 4x4 or NULL; Jgeom_out [B,36] row-major 6x6 or NULL. */
int uvs_ur10_fk(int64_t B, const double* q, double* T06_out, double* Jgeom_out, void* stream);

/* Pinhole projection of the nf disc centres seen from DH frame 6 (replaces the
 * renderer + utils.py:124-166).  targets_host [nf*3] HOST; f_out [B,m];
 * depth_out [B,nf] or NULL. */
int uvs_project(int32_t nf, int64_t B, const double* q, const double* targets_host,
                double* f_out, double* depth_out, void* stream);

/* Initial guess X0 = vec(J_image kron(I2,R06^T) J_geom) — experiment.py:94-114.
 * X0_out [B,N]; f_out [B,m] (the noise-free features at q) or NULL. */
int uvs_initial_guess(int32_t nf, int64_t B, const double* q, const double* targets_host,
                      double* X0_out, double* f_out, void* stream);

/* Counter-based noise streams — noise.py:81-207 on Philox4x32-10
 * (key = seed, counter = trial, frame, coordinate, draw).  out [B,T,m];
 * trial b of the batch is global trial trial0+b. */
int uvs_noise_fill(const uvs_noise_params* np, uint64_t trial0, int64_t B, int32_t T, int32_t m,
                   double* out, void* stream);

/* Start poses of main.py:129-134 from Philox (key = pose_seed).  q_out [B,6]. */
int uvs_start_poses(const double* q_start_host, uint64_t pose_seed, int32_t change_q_start,
                    uint64_t trial0, int64_t B, double* q_out, void* stream);

/* B closed-loop trials, T frames each, in one launch — experiment.py:48-359
 * with the simulator replaced by the kinematic UR10 + pinhole model. */
int uvs_closed_loop(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B,
                    const uvs_closed_loop_io* io, void* stream);

/* B trials of the tests/-style camera-twist loop (see uvs_twist_cfg).  cfg->filt.method may be
 * UVS_METHOD_ANALYTICAL (tests/ibvs_*.py) or any filter; general covariance (dense family) only.
 * io as for uvs_closed_loop; stats use t_k = (k+1) dt; frames[b] < T when the threshold stopped the trial. */
int uvs_twist_loop(const uvs_closed_loop_cfg* cfg, const uvs_twist_cfg* tw, uint64_t trial0, int64_t B,
                   const uvs_closed_loop_io* io, void* stream);

/* Same with HOST result buffers (stats_host [B,3], status_host [B],
 * frames_host [B]): runs on the calling thread's cudaStreamPerThread, copies the
 * results back (a DMA when the buffers are page-locked) and synchronises.  The
 * device-side result scratch is owned by the library: one block per host thread and device
 * that only grows (32 B per trial) and is freed when that host thread exits. */
int uvs_closed_loop_host(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B,
                         double* stats_host, int32_t* status_host, int32_t* frames_host);

/* Trajectory-log streaming — the batch form of main.py:152-196, which appends every experiment's T frame rows to
 * results.csv: the [B,T,6+2m] log of a large job (53 GB for 10^6 trials of 300 frames) never exists in one piece.
 * The batch runs as launches of `chunk` trials; each chunk's log goes to one of two device buffers, is copied to one
 * of two page-locked host buffers on a second stream while the next chunk computes, and is handed to `sink` on the
 * calling thread (which may format CSV rows, write a binary file, ...) while the GPU works on.
 * sink(user, first_trial, n, log_host [n,T,6+2m], stats_host [n,3], status_host [n], frames_host [n]); the pointers are
 * valid only during the call; a non-zero return stops the stream (UVS_ERR_INVALID).  sink may be NULL (measurement).
 * Synchronous.  Optional host outputs: d2h_gbs = bytes copied / time the copy engine was busy (CUDA events on the copy
 * stream), wall_s = duration of the whole call.  cfg->filt / noise / sweep as for uvs_closed_loop. */
typedef int (*uvs_log_sink)(void* user, uint64_t first_trial, int64_t n, const double* log_host, const double* stats_host,
                            const int32_t* status_host, const int32_t* frames_host);
int uvs_closed_loop_stream(const uvs_closed_loop_cfg* cfg, uint64_t trial0, int64_t B, int64_t chunk,
                           uvs_log_sink sink, void* user, double* d2h_gbs, double* wall_s);

/* Per-noise-level sums for results/plot_errorbar.m:87-94: for level c and
 * statistic s in {ISE,IAE,ITAE}: sums[c][s] = {n_ok, sum x, sum x^2} over
 * SUCCESS trials, plus n_fail[c].  level = (trial0 + b * trial_stride) / sweep_epoch (clamped; stride 0 = 1).
 * sums [n_levels,3,3] and n_fail [n_levels] are ACCUMULATED into (zero them
 * first); they are what is all-reduced across GPUs.  1 <= n_levels <= 512. */
int uvs_reduce_stats(int64_t B, uint64_t trial0, int32_t trial_stride, int32_t sweep_epoch, int32_t n_levels,
                     const double* stats, const int32_t* status, double* sums, double* n_fail,
                     void* stream);

/* Measured FP64 FMA peak of the device (roofline denominator): runs a
 * register-resident DFMA kernel for `iters` iterations on every SM and writes
 * the achieved TFLOP/s to *tflops_host.  Synchronous. */
int uvs_fp64_peak(int32_t iters, double* tflops_host, double* ms_host);

#ifdef __cplusplus
}
#endif
#endif /* UVS_B200_H */
