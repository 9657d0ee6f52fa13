// ur10.cuh — kinematic UR10 + pinhole camera (device side).
// DH chain and geometric Jacobian follow /root/reference/ur10_simulation.py:97-139,204-211;
// the camera frame is DH frame 6 and the pixel model is SURVEY.md Appendix A.
#pragma once
#include "common.cuh"

namespace uvs {

// 3x4 homogeneous transform (last row implicit 0 0 0 1), row-major
struct Tf {
    double r[3][3];
    double p[3];
};

// T <- T * dh(theta, d, a, alpha)   with   dh = Rz(theta) Tz(d) Tx(a) Rx(alpha)
// (ur10_simulation.py:204-211), given s = sin(theta), c = cos(theta).  ca/sa are compile-time so that
// zero terms vanish.
template <int CA, int SA>
UVS_DEV void dh_mul(Tf& T, double s, double c, double d, double a) {
    constexpr double ca = (double)CA, sa = (double)SA;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double t0 = T.r[i][0], t1 = T.r[i][1], t2 = T.r[i][2];
        const double x = c * t0 + s * t1;          // column 0 of T*A
        const double y = c * t1 - s * t0;          // (-s t0 + c t1)
        T.r[i][0] = x;
        T.r[i][1] = ca * y + sa * t2;
        T.r[i][2] = ca * t2 - sa * y;
        T.p[i] = T.p[i] + a * x + d * t2;
    }
}

struct UR10 {
    Tf T[6];       // T_0_1 .. T_0_6
};

// joint angle -> DH theta (ur10_simulation.py:100-105)
UVS_DEV double dh_theta(int j, double qj) {
    return (j == 1 || j == 3) ? qj - M_PI / 2 : ((j == 5) ? qj + M_PI : qj);
}

// the DH chain given sin/cos of the six thetas
UVS_DEV void ur10_chain(const double (&s)[6], const double (&c)[6], Tf& T6, UR10* all = nullptr) {
    Tf T;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) T.r[i][j] = (i == j) ? 1.0 : 0.0;
        T.p[i] = 0.0;
    }
    dh_mul<0, -1>(T, s[0], c[0], 0.128, 0.0);    if (all) all->T[0] = T;
    dh_mul<1, 0>(T, s[1], c[1], 0.0, 0.6127);    if (all) all->T[1] = T;
    dh_mul<1, 0>(T, s[2], c[2], 0.0, 0.5716);    if (all) all->T[2] = T;
    dh_mul<0, -1>(T, s[3], c[3], 0.1639, 0.0);   if (all) all->T[3] = T;
    dh_mul<0, 1>(T, s[4], c[4], 0.1157, 0.0);    if (all) all->T[4] = T;
    dh_mul<1, 0>(T, s[5], c[5], 0.0922, 0.0);    if (all) all->T[5] = T;
    T6 = T;
}

// p <- dh(theta, d, a, alpha)^-1 p : the point seen from the next link frame.  With A = [R | t],
// R = Rz(theta) Rx(alpha), t = (a c, a s, d):  A^-1 p = R^T (p - t) = Rx(-alpha) (Rz(-theta) p - (a, 0, d)).
template <int CA, int SA>
UVS_DEV void dh_inv_point(double (&p)[3], double s, double c, double d, double a) {
    constexpr double ca = (double)CA, sa = (double)SA;
    const double xr = c * p[0] + s * p[1] - a;
    const double yr = c * p[1] - s * p[0];
    const double zr = p[2] - d;
    p[0] = xr;
    p[1] = ca * yr + sa * zr;
    p[2] = ca * zr - sa * yr;
}

// World point -> camera (DH frame 6) coordinates by walking the chain backwards: 6 x ~6 flops instead of
// the ~110 of the forward matrix chain + R^T (p - t).  Same value up to rounding (a few ulp).
UVS_DEV void ur10_point_in_camera(const double (&s)[6], const double (&c)[6], const double* pw, double (&pc)[3]) {
    pc[0] = pw[0]; pc[1] = pw[1]; pc[2] = pw[2];
    dh_inv_point<0, -1>(pc, s[0], c[0], 0.128, 0.0);
    dh_inv_point<1, 0>(pc, s[1], c[1], 0.0, 0.6127);
    dh_inv_point<1, 0>(pc, s[2], c[2], 0.0, 0.5716);
    dh_inv_point<0, -1>(pc, s[3], c[3], 0.1639, 0.0);
    dh_inv_point<0, 1>(pc, s[4], c[4], 0.1157, 0.0);
    dh_inv_point<1, 0>(pc, s[5], c[5], 0.0922, 0.0);
}

// ---- scene plant (uvs_closed_loop_cfg.sensor_offset / disc_radius) ------------------------
// The reference measures a feature as the centroid of the pixels of a coloured disc (utils.py:73-78:
// 255 sum(X mask) / sum(mask) on the flipped 256 x 256 image).  Here the disc (radius r, lying in a plane z = const of
// the world, as in the scenes) is seen through the pinhole as the polygon through kDiscVerts projected rim points; its
// centroid (shoelace formula) is the feature.  With `clip` the polygon is first clipped to the image area
// [-0.5, 255.5]^2 (Sutherland-Hodgman): a partly visible disc yields a biased centroid, a disc with no visible part
// yields area 0 (the reference: 0/0 = NaN -> np.linalg.pinv raises -> FAIL).
constexpr int kDiscVerts = 16;
constexpr double kImageLo = -0.5, kImageHi = 255.5;

// one Sutherland-Hodgman pass against the half plane  sgn * (coord[axis] - bound) <= 0;  returns the new count
UVS_DEV int clip_half_plane(const double (*in)[2], int n, double (*out)[2], int axis, double bound, double sgn) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const double* a = in[i];
        const double* b = in[(i + 1 == n) ? 0 : i + 1];
        const double da = sgn * (a[axis] - bound), db = sgn * (b[axis] - bound);
        const bool ina = da <= 0.0, inb = db <= 0.0;
        if (ina) { out[m][0] = a[0]; out[m][1] = a[1]; ++m; }
        if (ina != inb) {
            const double t = da / (da - db);
            out[m][0] = a[0] + t * (b[0] - a[0]);
            out[m][1] = a[1] + t * (b[1] - a[1]);
            out[m][axis] = bound;
            ++m;
        }
    }
    return m;
}

// Centroid (u, v) of the image region of the disc centred at pw, given sin / cos of the DH angles.  off = sensor origin
// in DH frame 6.  Returns the signed-area magnitude of the (clipped) region in px^2: 0 = nothing visible.
// depth = z of the disc centre in the sensor frame.
UVS_DEV double disc_centroid(const double (&sn)[6], const double (&cs)[6], const double* pw, const double* off, double radius,
                             bool clip, double& u, double& v, double& depth) {
    double c0[3], c1[3], c2[3];
    const double px[3] = {pw[0] + radius, pw[1], pw[2]}, py[3] = {pw[0], pw[1] + radius, pw[2]};
    ur10_point_in_camera(sn, cs, pw, c0);
    ur10_point_in_camera(sn, cs, px, c1);
    ur10_point_in_camera(sn, cs, py, c2);
#pragma unroll
    for (int j = 0; j < 3; ++j) { c1[j] -= c0[j]; c2[j] -= c0[j]; c0[j] -= off[j]; }      // rim axes; centre in the sensor frame
    depth = c0[2];
    double poly[kDiscVerts + 4][2], tmp[kDiscVerts + 4][2];
    bool behind = false, outside = false;
    for (int i = 0; i < kDiscVerts; ++i) {
        double sa, ca;
        sincospi(2.0 * i / kDiscVerts, &sa, &ca);
        const double x = c0[0] + ca * c1[0] + sa * c2[0], y = c0[1] + ca * c1[1] + sa * c2[1], z = c0[2] + ca * c1[2] + sa * c2[2];
        behind = behind || !(z > kMinDepth);
        poly[i][0] = kPrincipal + kFocal * x / z;
        poly[i][1] = kPrincipal + kFocal * y / z;
        outside = outside || !(poly[i][0] >= kImageLo && poly[i][0] <= kImageHi && poly[i][1] >= kImageLo && poly[i][1] <= kImageHi);
    }
    if (behind) { u = v = NAN; return 0.0; }
    int n = kDiscVerts;
    const double (*P)[2] = poly;
    if (clip && outside) {
        n = clip_half_plane(poly, n, tmp, 0, kImageLo, -1.0);
        n = clip_half_plane(tmp, n, poly, 0, kImageHi, 1.0);
        n = clip_half_plane(poly, n, tmp, 1, kImageLo, -1.0);
        n = clip_half_plane(tmp, n, poly, 1, kImageHi, 1.0);
        if (n < 3) { u = v = NAN; return 0.0; }
    }
    // shoelace about the first vertex (keeps the cancellation small)
    const double x0 = P[0][0], y0 = P[0][1];
    double a2 = 0.0, cx = 0.0, cy = 0.0;
    for (int i = 1; i + 1 < n; ++i) {
        const double ax = P[i][0] - x0, ay = P[i][1] - y0, bx = P[i + 1][0] - x0, by = P[i + 1][1] - y0;
        const double cr = ax * by - bx * ay;
        a2 += cr;
        cx += (ax + bx) * cr;
        cy += (ay + by) * cr;
    }
    if (!(fabs(a2) > 1e-9)) { u = v = NAN; return 0.0; }
    u = x0 + cx / (3.0 * a2);
    v = y0 + cy / (3.0 * a2);
    return 0.5 * fabs(a2);
}

// Feature coordinate `axis` (0 = u, 1 = v) of the target at pw under the configured plant: the projected centre (seen
// from the offset sensor) or the disc centroid.  visible = false when the reference would measure NaN.
UVS_DEV double scene_feature(const double (&sn)[6], const double (&cs)[6], const double* pw, const uvs_closed_loop_cfg& cfg,
                             int axis, double& depth, bool& visible) {
    if (cfg.disc_radius > 0.0) {
        double u, v;
        const double area = disc_centroid(sn, cs, pw, cfg.sensor_offset, cfg.disc_radius, cfg.fov_check != 0, u, v, depth);
        visible = area > 0.0;
        return axis ? v : u;
    }
    double pc[3];
    ur10_point_in_camera(sn, cs, pw, pc);
#pragma unroll
    for (int j = 0; j < 3; ++j) pc[j] -= cfg.sensor_offset[j];
    depth = pc[2];
    const double f = kPrincipal + kFocal * (axis ? pc[1] : pc[0]) / pc[2];
    visible = (f >= 0.0 && f < kResolution && depth > kMinDepth);
    return f;
}

__host__ __device__ inline bool scene_plant(const uvs_closed_loop_cfg& cfg) {
    return cfg.disc_radius > 0.0 || cfg.sensor_offset[0] != 0.0 || cfg.sensor_offset[1] != 0.0 || cfg.sensor_offset[2] != 0.0 ||
           cfg.servo_substeps > 0;
}

// sin / cos of the six DH thetas for a group of GS lanes that all hold the same q: the evaluations are
// split over the lanes of the group and exchanged with shuffles; gl = lane index in the group.  Every
// lane of the warp must call.
template <int GS>
UVS_DEV void ur10_sincos_group(const double (&q)[6], int gl, double (&s)[6], double (&c)[6]) {
    constexpr int SLOTS = (6 + GS - 1) / GS;
    double ss[SLOTS], cs[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
        double th = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j)
            if (j / GS == t && j % GS == gl) th = dh_theta(j, q[j]);
        sincos(th, &ss[t], &cs[t]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        s[j] = __shfl_sync(UVS_FULL_MASK, ss[j / GS], j % GS, GS);
        c[j] = __shfl_sync(UVS_FULL_MASK, cs[j / GS], j % GS, GS);
    }
}

// pixel coordinate `axis` of a camera-frame point (SURVEY Appendix A pixel model)
UVS_DEV double pixel_of(const double (&pc)[3], int axis) {
    return kPrincipal + kFocal * (axis ? pc[1] : pc[0]) / pc[2];
}

// ur10_simulation.py:97-110 — every thread evaluates all six sincos
UVS_DEV void ur10_fk(const double (&q)[6], Tf& T6, UR10* all = nullptr) {
    double s[6], c[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) sincos(dh_theta(j, q[j]), &s[j], &c[j]);
    ur10_chain(s, c, T6, all);
}

// Same result for a group of GS lanes that all hold the same q (sincos split over the lanes).
template <int GS>
UVS_DEV void ur10_fk_group(const double (&q)[6], int gl, Tf& T6) {
    double s[6], c[6];
    ur10_sincos_group<GS>(q, gl, s, c);
    ur10_chain(s, c, T6);
}

// ur10_simulation.py:112-139 — J[row][col], rows = (v; w), cols = joints
UVS_DEV void ur10_jacobian(const UR10& k, double (&J)[6][6]) {
    const double* pe = k.T[5].p;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double z[3], p[3];
        if (i == 0) { z[0] = 0; z[1] = 0; z[2] = 1; p[0] = p[1] = p[2] = 0; }
        else {
            z[0] = k.T[i - 1].r[0][2]; z[1] = k.T[i - 1].r[1][2]; z[2] = k.T[i - 1].r[2][2];
            p[0] = k.T[i - 1].p[0]; p[1] = k.T[i - 1].p[1]; p[2] = k.T[i - 1].p[2];
        }
        const double dx = pe[0] - p[0], dy = pe[1] - p[1], dz = pe[2] - p[2];
        J[0][i] = z[1] * dz - z[2] * dy;
        J[1][i] = z[2] * dx - z[0] * dz;
        J[2][i] = z[0] * dy - z[1] * dx;
        J[3][i] = z[0]; J[4][i] = z[1]; J[5][i] = z[2];
    }
}

// point in the camera (DH-6) frame: pc = R^T (pw - t)
UVS_DEV void to_camera(const Tf& T, const double* pw, double (&pc)[3]) {
    const double dx = pw[0] - T.p[0], dy = pw[1] - T.p[1], dz = pw[2] - T.p[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) pc[j] = T.r[0][j] * dx + T.r[1][j] * dy + T.r[2][j] * dz;
}

// pixel coordinate `axis` (0 = u, 1 = v) of world point pw; depth returned through z
UVS_DEV double project_coord(const Tf& T, const double* pw, int axis, double& z) {
    double pc[3];
    to_camera(T, pw, pc);
    z = pc[2];
    return kPrincipal + kFocal * (axis ? pc[1] : pc[0]) / pc[2];
}

// ur10_simulation.py:188 — Euclidean camera-to-disc distance used as "Z"
UVS_DEV double camera_distance(const Tf& T, const double* pw) {
    const double dx = T.p[0] - pw[0], dy = T.p[1] - pw[1], dz = T.p[2] - pw[2];
    return sqrt(dx * dx + dy * dy + dz * dz);
}

// One row of X0 = J_image kron(I2, R^T) J_geom (experiment.py:94-114).
// coord = 2*feature + axis; u,v = raw pixel of the feature; Zc = camera_distance.
UVS_DEV void initial_guess_row(const UR10& k, const double (&Jg)[6][6], int coord, double u, double v,
                               double Zc, double (&row)[6]) {
    const double f = kFocal;
    double ji[6];
    if ((coord & 1) == 0) {
        ji[0] = -f / Zc; ji[1] = 0.0; ji[2] = u / Zc; ji[3] = u * v / f; ji[4] = -(f * f + u * u) / f; ji[5] = v;
    } else {
        ji[0] = 0.0; ji[1] = -f / Zc; ji[2] = v / Zc; ji[3] = (f * f + v * v) / f; ji[4] = -u * v / f; ji[5] = -u;
    }
    // w = ji * kron(I2, R^T): w[0:3] = ji[0:3] R^T, w[3:6] = ji[3:6] R^T
    const Tf& T = k.T[5];
    double w[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        w[c] = ji[0] * T.r[c][0] + ji[1] * T.r[c][1] + ji[2] * T.r[c][2];
        w[3 + c] = ji[3] * T.r[c][0] + ji[4] * T.r[c][1] + ji[5] * T.r[c][2];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc += w[c] * Jg[c][j];
        row[j] = acc;
    }
}

}  // namespace uvs
