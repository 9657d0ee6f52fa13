// philox.cuh — Philox4x32-10 counter-based streams and the noise models of
// /root/reference/noise.py:120-207 (device side).  Counter layout documented in
// oracle/noise_model.py; this file and that one must stay in lock step.
#pragma once
#include "common.cuh"

namespace uvs {

enum { PURPOSE_NOISE = 0, PURPOSE_POSE = 1, PURPOSE_X0 = 2 };

UVS_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t (&out)[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// ((x >> 12) + 0.5) * 2^-52 : exact, strictly inside (0,1)
UVS_DEV double open52(uint32_t lo, uint32_t hi) {
    const uint64_t x = ((uint64_t)hi << 32) | (uint64_t)lo;
    return ((double)(x >> 12) + 0.5) * 2.220446049250313e-16;
}

UVS_DEV void uniforms(uint64_t seed, uint64_t trial, uint32_t frame, uint32_t coord, uint32_t draw,
                      uint32_t purpose, double& ua, double& ub) {
    uint32_t r[4];
    philox4x32_10((uint32_t)trial, (uint32_t)(trial >> 32), frame, (purpose << 16) | (draw << 8) | coord,
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    ua = open52(r[0], r[1]);
    ub = open52(r[2], r[3]);
}

UVS_DEV double box_muller(double ua, double ub) {
    return sqrt(-2.0 * log(ua)) * cos(2.0 * M_PI * ub);
}

// One fresh draw for (trial, frame, coord) — noise.py:120-207.
UVS_DEV double noise_sample(const uvs_noise_params& np, double sweep_rho, double sweep_alpha,
                            uint64_t trial, uint32_t frame, uint32_t coord) {
    double ua, ub;
    uniforms(np.seed, trial, frame, coord, 0, PURPOSE_NOISE, ua, ub);
    switch (np.type) {
    case UVS_NOISE_UNIFORM:                                  // noise.py:120-123
        return ua;
    case UVS_NOISE_WHITE:                                    // noise.py:125-128
        return np.std * box_muller(ua, ub);
    case UVS_NOISE_GAUSSIAN_MIXTURE:                         // noise.py:130-138
    case UVS_NOISE_GAUSSIAN_BIMODAL: {                       // noise.py:140-150
        const double n = box_muller(ua, ub);
        double sel, dummy;
        uniforms(np.seed, trial, frame, coord, 1, PURPOSE_NOISE, sel, dummy);
        const double rho = sweep_rho;
        double loc;
        if (np.type == UVS_NOISE_GAUSSIAN_MIXTURE) loc = (sel > rho) ? 0.0 : np.mean;
        else loc = (sel > rho) ? 0.0 : ((sel > rho / 2) ? np.mean : -np.mean);
        return loc + np.std * n;
    }
    case UVS_NOISE_ALPHA_STABLE: {                           // noise.py:152-207
        const double alpha = sweep_alpha, beta = np.beta;
        double value;
        if (alpha == 2.0) {
            value = sqrt(2.0) * box_muller(ua, ub);
        } else if (alpha == 1.0 && beta == 0.0) {
            value = tan(-M_PI / 2 + M_PI * ua);
        } else if (alpha == 0.5 && fabs(beta) == 1.0) {
            const double n = box_muller(ua, ub);
            value = beta / (n * n);
        } else {
            const double V = -M_PI / 2 + M_PI * ua;
            const double W = -log(ub);
            if (beta == 0.0) {
                // sin(aV) / cos(V)^(1/a) * (cos((1-a)V) / W)^((1-a)/a), powers through exp/log (all
                // bases are positive here): ~2x cheaper than two pow() calls, relative error < 1e-14
                // The three angles lie in (-pi, pi): the bounded sincos (common.cuh) replaces sin / cos / cos, and
                // log(cos((1-a)V)) - log(W) is one logarithm of the branch-free quotient (both positive, tame).
                const double ia = 1.0 / alpha;
                double s_av, c_av, s_bv, c_bv, s_v, c_v;
                sincos_bounded(alpha * V, &s_av, &c_av);
                sincos_bounded(V * (1.0 - alpha), &s_bv, &c_bv);
                sincos_bounded(V, &s_v, &c_v);
                value = s_av * exp((1.0 - alpha) * ia * log(fast_div(c_bv, W)) - ia * log(c_v));
            } else if (alpha != 1.0) {
                const double constant = beta * tan(M_PI * alpha / 2);
                const double Bc = atan(constant);
                const double S = pow(1.0 + constant * constant, 1.0 / (2.0 * alpha));
                value = S * sin(alpha * V + Bc) / pow(cos(V), 1.0 / alpha) *
                        pow(cos((1.0 - alpha) * V - Bc) / W, (1.0 - alpha) / alpha);
            } else {
                const double sclshftV = M_PI / 2 + beta * V;
                value = 2.0 / M_PI * (sclshftV * tan(V) - beta * log((W * cos(V)) / sclshftV));
            }
        }
        if (alpha == 1.0)
            return np.gamma * value + (2.0 / M_PI) * beta * np.gamma * log(np.gamma) + np.delta;
        return np.gamma * value + np.delta;
    }
    default:
        return 0.0;
    }
}

// Per-coordinate state of the pair-hold logic (noise.py:81-118).  Both lanes of
// a feature pair keep identical counters; `partner_xor` exchanges the outlier bit.
struct HoldState {
    double cur = 0.0;
    int cnt = 0, cnt_max = 0;
};

}  // namespace uvs
