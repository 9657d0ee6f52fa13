// kernel_args.cuh — argument blocks and per-trial helpers shared by the dense and the
// block-diagonal kernel families.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "ur10.cuh"

namespace uvs {

struct ReplayArgs {
    uvs_filter_params fp;
    long long B;
    int T, k_max, direct, p_layout;
    double sigma;
    const double *f, *dq, *f_init;     // replay inputs
    const double *h, *Z;               // direct (single step) inputs
    double *X, *P;                     // in/out state
    double *X_out, *P_out, *kappa_out, *K_out;
    int32_t* flags_out;
};

struct ClosedLoopArgs {
    uvs_closed_loop_cfg cfg;
    uvs_closed_loop_io io;
    unsigned long long trial0;
    long long B;
    unsigned long long stride;         // global trial of batch entry b = trial0 + b * stride (cfg.trial_stride, >= 1)
#ifdef UVS_PHASE_CLOCKS
    unsigned long long* phase_clk;     // profiling build: device array of UVS_NPHASE totals
#endif
};

// noise of this frame for coordinate `lane` with the pair hold of noise.py:81-118
UVS_DEV double next_noise(HoldState& hs, const uvs_noise_params& np, double rho, double alpha, uint64_t trial,
                          uint32_t frame, int lane, int M) {
    bool redraw = false;
    double fresh = 0.0;
    if (lane < M) {
        if (hs.cnt >= hs.cnt_max) {
            hs.cnt = 0;
            redraw = true;
            fresh = noise_sample(np, rho, alpha, trial, frame, (uint32_t)lane);
        } else {
            hs.cnt += 1;
        }
    }
    const bool outl = redraw && (fabs(fresh) > 20.0);
    // the shuffle must be executed by every lane: no short-circuit in front of it
    const bool partner = __shfl_xor_sync(UVS_FULL_MASK, outl, 1) != 0;
    const bool pair = outl || partner;
    if (redraw) {
        hs.cnt_max = pair ? np.hold_cnt : 0;
        hs.cur = fresh;
    }
    return hs.cur;
}

// Same hold logic when this frame's fresh draw was computed ahead of time (counter-based streams make
// the draw a pure function of (trial, frame, coordinate), so it can be evaluated by an otherwise idle
// lane several frames early).  Every lane of the warp must call.
UVS_DEV double next_noise_prefetched(HoldState& hs, int hold_cnt, double fresh, int lane, int M) {
    bool redraw = false;
    if (lane < M) {
        if (hs.cnt >= hs.cnt_max) { hs.cnt = 0; redraw = true; }
        else hs.cnt += 1;
    }
    const bool outl = redraw && (fabs(fresh) > 20.0);
    const bool partner = __shfl_xor_sync(UVS_FULL_MASK, outl, 1) != 0;
    if (redraw) {
        hs.cnt_max = (outl || partner) ? hold_cnt : 0;
        hs.cur = fresh;
    }
    return hs.cur;
}

// joint velocity actually executed for the commanded rate dq (uvs_closed_loop_cfg.v_max / a_max; 0 = unlimited)
UVS_DEV double servo_rate(const uvs_closed_loop_cfg& cfg, double dq, double v_prev) {
    double v = (cfg.servo_gain > 0.0) ? cfg.servo_gain * dq : dq;
    if (cfg.a_max > 0.0) {
        const double dv = cfg.a_max * cfg.dt;
        v = fmin(fmax(v, v_prev - dv), v_prev + dv);
    }
    if (cfg.v_max > 0.0) v = fmin(fmax(v, -cfg.v_max), cfg.v_max);
    return v;
}

// One frame of the sub-stepped position servo (uvs_closed_loop_cfg.servo_substeps / servo_kp): the joint runs a P
// controller towards q_target = q + dq dt (experiment.py:320) at the dynamics step h = dt / substeps, under the scenes'
// acceleration / velocity limits.  q and v_prev (the joint rate of the last substep) are updated in place.
UVS_DEV void servo_substepped(const uvs_closed_loop_cfg& cfg, double dq, double& q, double& v_prev) {
    const double target = q + dq * cfg.dt;
    const double h = cfg.dt / (double)cfg.servo_substeps;
    const double dv = cfg.a_max * h;
    for (int s = 0; s < cfg.servo_substeps; ++s) {
        double v = cfg.servo_kp * (target - q) / h;
        if (cfg.a_max > 0.0) v = fmin(fmax(v, v_prev - dv), v_prev + dv);
        if (cfg.v_max > 0.0) v = fmin(fmax(v, -cfg.v_max), cfg.v_max);
        q += v * h;
        v_prev = v;
    }
}

UVS_DEV void start_pose(const uvs_closed_loop_cfg& cfg, uint64_t trial, double (&q)[6]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) q[j] = cfg.q_start[j];
    if (cfg.change_q_start) {                                   // main.py:129-134
        double r0, r1, dummy;
        uniforms(cfg.pose_seed, trial, 0, 0, 0, PURPOSE_POSE, r0, dummy);
        uniforms(cfg.pose_seed, trial, 0, 1, 0, PURPOSE_POSE, r1, dummy);
        q[0] = cfg.q_start[0] + 2 * (r0 - 1) * (M_PI / 18);
        q[1] = cfg.q_start[1] + 2 * (r1 - 1) * (M_PI / 9);
    }
}

}  // namespace uvs
