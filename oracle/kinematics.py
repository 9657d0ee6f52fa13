"""Oracle (test infrastructure): kinematic UR10 + pinhole camera model.

Follows /root/reference/ur10_simulation.py:97-154,204-211 (DH chain, geometric
Jacobian, camera rotation/position), experiment.py:94-114 (initial guess of the
feature Jacobian) and the scene constants recovered from scenarios/*.ttt
(SURVEY.md Appendix A).  Plain NumPy, one pose at a time.
"""
import numpy as np

RESOLUTION = 256
PERSPECTIVE_ANGLE_DEG = 65.0
# experiment.py:97  focal = W / (2 tan(angle/2))
FOCAL = RESOLUTION / (2.0 * np.tan(0.5 * np.deg2rad(PERSPECTIVE_ANGLE_DEG)))
PRINCIPAL = 127.5

# world positions of the disc centres, feature order = red, green, blue, pink
# (utils.py:73-78,143-144; positions from the .ttt scenes)
TARGETS = {
    1: np.array([[0.650, 0.000, 0.012]]),
    3: np.array([[0.575, -0.075, 0.012],
                 [0.650, 0.000, 0.012],
                 [0.575, 0.075, 0.012]]),
    4: np.array([[0.625, 0.100, 0.012],
                 [0.700, 0.175, 0.012],
                 [0.625, 0.250, 0.012],
                 [0.550, 0.175, 0.012]]),
}

# ur10_simulation.py:100-105  (theta offset, d, a, alpha)
DH_TABLE = (
    (0.0,        0.128,  0.0,    -np.pi / 2),
    (-np.pi / 2, 0.0,    0.6127, 0.0),
    (0.0,        0.0,    0.5716, 0.0),
    (-np.pi / 2, 0.1639, 0.0,    -np.pi / 2),
    (0.0,        0.1157, 0.0,    np.pi / 2),
    (np.pi,      0.0922, 0.0,    0.0),
)


def _trotz(theta, d):
    # ur10_simulation.py:207-208
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0, 0.0], [s, c, 0.0, 0.0],
                     [0.0, 0.0, 1.0, d], [0.0, 0.0, 0.0, 1.0]])


def _trotx(alpha, a):
    # ur10_simulation.py:204-205
    c, s = np.cos(alpha), np.sin(alpha)
    return np.array([[1.0, 0.0, 0.0, a], [0.0, c, -s, 0.0],
                     [0.0, s, c, 0.0], [0.0, 0.0, 0.0, 1.0]])


def dh(theta, d, a, alpha):
    # ur10_simulation.py:210-211
    return _trotz(theta, d) @ _trotx(alpha, a)


def fkine_all(q):
    """T_0_1 .. T_0_6 (ur10_simulation.py:97-110)."""
    T = np.eye(4)
    out = []
    for i, (off, d, a, alpha) in enumerate(DH_TABLE):
        T = T @ dh(q[i] + off, d, a, alpha)
        out.append(T)
    return out


def fkine(q):
    return fkine_all(q)[-1]


def jacobian(q):
    """Geometric Jacobian, ur10_simulation.py:112-139."""
    return jacobian_from_transforms(fkine_all(q))


def jacobian_from_transforms(Ts):
    """The same from a given list T_0_1 .. T_0_6 — UR10Simulation.jacobian() works on its CACHED transforms, which are
    only refreshed by fkine(recalculate=True) (ur10_simulation.py:97-115)."""
    p_e = Ts[5][0:3, 3]
    z = [np.array([0.0, 0.0, 1.0])] + [Ts[i][0:3, 2] for i in range(5)]
    p = [np.zeros(3)] + [Ts[i][0:3, 3] for i in range(5)]
    cols = [np.concatenate([np.cross(z[i], p_e - p[i]), z[i]]) for i in range(6)]
    return np.stack(cols, axis=1)


def project(q, targets):
    """Pixel coordinates (u, v) of the disc centres seen from DH frame 6
    (camera frame == DH-6, SURVEY.md Appendix A 'Pixel model')."""
    T = fkine(q)
    R, t = T[0:3, 0:3], T[0:3, 3]
    f = np.zeros(2 * len(targets))
    for i, pw in enumerate(targets):
        pc = R.T @ (pw - t)
        f[2 * i] = PRINCIPAL + FOCAL * pc[0] / pc[2]
        f[2 * i + 1] = PRINCIPAL + FOCAL * pc[1] / pc[2]
    return f


def depth(q, targets):
    """Z of each disc centre in the camera frame (for the field-of-view rule)."""
    T = fkine(q)
    R, t = T[0:3, 0:3], T[0:3, 3]
    return np.array([(R.T @ (pw - t))[2] for pw in targets])


def compute_Z(q, targets):
    """ur10_simulation.py:170-195 — Euclidean distance camera -> disc centre."""
    t = fkine(q)[0:3, 3]
    return np.array([np.linalg.norm(t - pw) for pw in targets])


def initial_guess(q, targets, f_measured=None):
    """X0 = vec_rowmajor(J_image kron(I2, R06^T) J_geom), experiment.py:94-114.

    The interaction rows use the RAW pixel (u, v) and the Euclidean distance,
    exactly as the reference does.  f_measured: what the detector reports at q when the plant is not the ideal
    point-feature pinhole (scene plant); the depth stays computeZ's DH-6-origin distance.
    """
    f = project(q, targets) if f_measured is None else np.array(f_measured, dtype=float)
    Zc = compute_Z(q, targets)
    m = len(f)
    J_image = np.zeros((m, 6))
    focal = FOCAL
    for i in range(m // 2):
        u, v = f[2 * i], f[2 * i + 1]
        J_image[2 * i, 0] = -focal / Zc[i]
        J_image[2 * i + 1, 1] = -focal / Zc[i]
        J_image[2 * i, 2] = u / Zc[i]
        J_image[2 * i + 1, 2] = v / Zc[i]
        J_image[2 * i, 3] = u * v / focal
        J_image[2 * i + 1, 3] = (focal ** 2 + v ** 2) / focal
        J_image[2 * i, 4] = -(focal ** 2 + u ** 2) / focal
        J_image[2 * i + 1, 4] = -u * v / focal
        J_image[2 * i, 5] = v
        J_image[2 * i + 1, 5] = -u
    R = fkine(q)[0:3, 0:3]
    J_feature = J_image @ np.kron(np.eye(2), R.T) @ jacobian(q)
    return J_feature.reshape(m * 6), f


def camera_pose(q):
    """Sensor position + roll/pitch/yaw (sensor = DH-6 . Rz(pi); euler as
    utils.py:174-179 applied to the sensor's wxyz quaternion)."""
    T = fkine(q)
    Rs = T[0:3, 0:3] @ np.diag([-1.0, -1.0, 1.0])
    pose = np.zeros(6)
    pose[0:3] = T[0:3, 3]
    pose[3] = np.arctan2(Rs[2, 1], Rs[2, 2])
    pose[4] = np.arcsin(np.clip(-Rs[2, 0], -1.0, 1.0))
    pose[5] = np.arctan2(Rs[1, 0], Rs[0, 0])
    return pose


# ---- scene plant (include/uvs_b200.h: sensor_offset / disc_radius) ------------------------
DISC_VERTS = 16
IMAGE_LO, IMAGE_HI = -0.5, 255.5
DISC_RADIUS_SCENE = 0.0375            # scenarios/*.ttt: disc primitives of 0.075 m diameter (tag Ppr of the *_disc shapes)


def _clip_half_plane(poly, axis, bound, sgn):
    out = []
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        da, db = sgn * (a[axis] - bound), sgn * (b[axis] - bound)
        ina, inb = da <= 0.0, db <= 0.0
        if ina:
            out.append([a[0], a[1]])
        if ina != inb:
            t = da / (da - db)
            p = [a[0] + t * (b[0] - a[0]), a[1] + t * (b[1] - a[1])]
            p[axis] = bound
            out.append(p)
    return out


def disc_centroid(q, pw, offset=(0.0, 0.0, 0.0), radius=DISC_RADIUS_SCENE, clip=True):
    """(u, v, area, depth): centroid of the image region of the disc centred at world point pw (radius `radius`, lying
    in a plane z = const) seen by the sensor mounted at `offset` in DH frame 6 — the quantity utils.py:73-78 measures
    (255 sum(X mask) / sum(mask)).  Same construction as csrc/ur10.cuh disc_centroid: DISC_VERTS-gon through the projected
    rim, optional Sutherland-Hodgman clip to the image, shoelace centroid.  area = 0 (u = v = nan): nothing visible."""
    T = fkine(q)
    R, t = T[0:3, 0:3], T[0:3, 3]
    off = np.asarray(offset, dtype=float)
    c0 = R.T @ (np.asarray(pw, dtype=float) - t)
    c1 = R.T @ (np.asarray(pw, dtype=float) + np.array([radius, 0.0, 0.0]) - t) - c0
    c2 = R.T @ (np.asarray(pw, dtype=float) + np.array([0.0, radius, 0.0]) - t) - c0
    c0 = c0 - off
    poly = []
    for i in range(DISC_VERTS):
        ang = 2.0 * np.pi * i / DISC_VERTS
        p = c0 + np.cos(ang) * c1 + np.sin(ang) * c2
        if not p[2] > 0.01:
            return np.nan, np.nan, 0.0, c0[2]
        poly.append([PRINCIPAL + FOCAL * p[0] / p[2], PRINCIPAL + FOCAL * p[1] / p[2]])
    outside = any(not (IMAGE_LO <= x <= IMAGE_HI and IMAGE_LO <= y <= IMAGE_HI) for x, y in poly)
    if clip and outside:
        poly = _clip_half_plane(poly, 0, IMAGE_LO, -1.0)
        poly = _clip_half_plane(poly, 0, IMAGE_HI, 1.0)
        poly = _clip_half_plane(poly, 1, IMAGE_LO, -1.0)
        poly = _clip_half_plane(poly, 1, IMAGE_HI, 1.0)
        if len(poly) < 3:
            return np.nan, np.nan, 0.0, c0[2]
    x0, y0 = poly[0]
    a2 = cx = cy = 0.0
    for i in range(1, len(poly) - 1):
        ax, ay = poly[i][0] - x0, poly[i][1] - y0
        bx, by = poly[i + 1][0] - x0, poly[i + 1][1] - y0
        cr = ax * by - bx * ay
        a2 += cr
        cx += (ax + bx) * cr
        cy += (ay + by) * cr
    if not abs(a2) > 1e-9:
        return np.nan, np.nan, 0.0, c0[2]
    return x0 + cx / (3.0 * a2), y0 + cy / (3.0 * a2), 0.5 * abs(a2), c0[2]


def scene_features(q, targets, offset=(0.0, 0.0, 0.0), radius=0.0, clip=False):
    """f[m] and visible[m] under the scene plant: projected centres seen from the offset sensor (radius = 0) or disc
    centroids (csrc/ur10.cuh scene_feature)."""
    m = 2 * len(targets)
    f = np.zeros(m)
    vis = np.ones(m, dtype=bool)
    off = np.asarray(offset, dtype=float)
    T = fkine(q)
    R, t = T[0:3, 0:3], T[0:3, 3]
    for i, pw in enumerate(targets):
        if radius > 0.0:
            u, v, area, _ = disc_centroid(q, pw, off, radius, clip)
            f[2 * i], f[2 * i + 1] = u, v
            vis[2 * i] = vis[2 * i + 1] = area > 0.0
        else:
            pc = R.T @ (pw - t) - off
            f[2 * i] = PRINCIPAL + FOCAL * pc[0] / pc[2]
            f[2 * i + 1] = PRINCIPAL + FOCAL * pc[1] / pc[2]
            for a in (0, 1):
                vis[2 * i + a] = (0.0 <= f[2 * i + a] < RESOLUTION) and pc[2] > 0.01
    return f, vis


def rasterized_disc_centroid(q, pw, offset=(0.0, 0.0, 0.0), radius=DISC_RADIUS_SCENE):
    """Brute force, the way the reference measures it (utils.py:73-78): every pixel (col j, row i) of the 256 x 256 image
    whose centre ray hits the disc belongs to the mask; f = (sum(j mask) / sum(mask), sum(i mask) / sum(mask)).
    Used by the CPU tests to pin the polygon model to the detector's semantics."""
    T = fkine(q)
    R, t = T[0:3, 0:3], T[0:3, 3]
    cam = t + R @ np.asarray(offset, dtype=float)             # sensor origin in the world
    jj, ii = np.meshgrid(np.arange(RESOLUTION), np.arange(RESOLUTION))
    d_cam = np.stack([(jj - PRINCIPAL) / FOCAL, (ii - PRINCIPAL) / FOCAL, np.ones_like(jj, dtype=float)], axis=-1)
    d_w = d_cam @ R.T                                          # ray directions in the world
    s = (pw[2] - cam[2]) / d_w[..., 2]                         # intersection with the plane z = pw[2]
    hit = cam[None, None, :] + s[..., None] * d_w
    mask = (s > 0.01) & ((hit[..., 0] - pw[0]) ** 2 + (hit[..., 1] - pw[1]) ** 2 <= radius ** 2)
    n = mask.sum()
    if n == 0:
        return np.nan, np.nan, 0
    return float((jj * mask).sum() / n), float((ii * mask).sum() / n), int(n)
