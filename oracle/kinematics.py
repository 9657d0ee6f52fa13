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


def initial_guess(q, targets):
    """X0 = vec_rowmajor(J_image kron(I2, R06^T) J_geom), experiment.py:94-114.

    The interaction rows use the RAW pixel (u, v) and the Euclidean distance,
    exactly as the reference does.
    """
    f = project(q, targets)
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
