"""numpy (float64) restatement of the quaternion / 6D-rotation helpers on the hot path.

Follows ``src/wear_mocap_ape/utility/transformations.py`` (``ts`` below); quaternions are ``[w, x, y, z]``.
All functions take arrays whose LAST axis is the component axis, so they serve single rows and columns of
rows alike (the reference branches on ``len(shape)`` for the same purpose).  Test infrastructure only.
"""
import numpy as np

_CONJ = np.array([1.0, -1.0, -1.0, -1.0])


def hamilton(a, b):
    """ts:129-149 - Hamilton product, component formulas of ts:141-144."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bw - ax * bx - ay * by - az * bz,
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
    ], axis=-1)


def rotate(q, v):
    """ts:83-126 - ``q (x) [0, v] (x) conj(q)`` WITHOUT normalising ``q`` (ts:105, ts:113/121)."""
    q, v = np.asarray(q, dtype=np.float64), np.asarray(v, dtype=np.float64)
    p = np.concatenate([np.zeros(v.shape[:-1] + (1,)), v], axis=-1)
    return hamilton(hamilton(q, p), q * _CONJ)[..., 1:]


def invert(q):
    """ts:244-254 - conjugate divided by the squared norm."""
    q = np.asarray(q, dtype=np.float64)
    return q * _CONJ / np.sum(np.square(q), axis=-1, keepdims=True)


def euler_to_quat(e):
    """ts:152-174 - roll/pitch/yaw half-angle products."""
    e = np.asarray(e, dtype=np.float64)
    cr, sr = np.cos(e[..., 0] * 0.5), np.sin(e[..., 0] * 0.5)
    cp, sp = np.cos(e[..., 1] * 0.5), np.sin(e[..., 1] * 0.5)
    cy, sy = np.cos(e[..., 2] * 0.5), np.sin(e[..., 2] * 0.5)
    return np.stack([
        cr * cp * cy + sr * sp * sy,
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
    ], axis=-1)


def android_to_global_no_north(q):
    """ts:225-229 - axis swap ``[-w, x, z, y]``."""
    q = np.asarray(q, dtype=np.float64)
    return np.stack([-q[..., 0], q[..., 1], q[..., 3], q[..., 2]], axis=-1)


def android_to_global(q, north):
    """ts:232-241 - ``north (x) swap(q)``."""
    return hamilton(north, android_to_global_no_north(q))


def y_rot_of(q):
    """ts:200-207 - rotate the forward vector [0,0,1] and take ``atan2(x, z)``."""
    pp = rotate(q, np.array([0.0, 0.0, 1.0]))
    return np.arctan2(pp[..., 0], pp[..., 2])


def north_quat(sw_fwd):
    """watch_only.py:67-69 - ``euler_to_quat([0, -y_rot, 0])`` of the swapped calibration quaternion."""
    y = y_rot_of(android_to_global_no_north(sw_fwd))
    z = np.zeros_like(y)
    return euler_to_quat(np.stack([z, -y, z], axis=-1))


def north_quat_left_arm(sw_fwd):
    """ts:182-197 - the north quaternion pre-multiplied by the left-hand calibration rotation."""
    return hamilton(np.array([0.7071068, 0.0, -0.7071068, 0.0]), north_quat(sw_fwd))


def quat_to_rot9(q):
    """ts:481-518 - transforms3d-style matrix with ``s = 2/|q|^2``; identity below eps (ts:492-493)."""
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    nq = w * w + x * x + y * y + z * z
    tiny = nq < np.finfo(np.float64).eps
    s = 2.0 / np.where(tiny, 1.0, nq)
    _x, _y, _z = x * s, y * s, z * s
    wx, wy, wz = w * _x, w * _y, w * _z
    xx, xy, xz = x * _x, x * _y, x * _z
    yy, yz, zz = y * _y, y * _z, z * _z
    m = np.stack([
        1.0 - (yy + zz), xy - wz, xz + wy,
        xy + wz, 1.0 - (xx + zz), yz - wx,
        xz - wy, yz + wx, 1.0 - (xx + yy),
    ], axis=-1)
    eye = np.array([1.0, 0, 0, 0, 1.0, 0, 0, 0, 1.0])
    return np.where(tiny[..., None], eye, m)


def quat_to_six(q):
    """ts:476-478 with ts:587-599 - first two matrix columns, row-major ``[r11,r12,r21,r22,r31,r32]``."""
    m = quat_to_rot9(q)
    return m[..., [0, 1, 3, 4, 6, 7]]


def six_to_rot9(six):
    """ts:602-637 - Gram-Schmidt of the two 3-vectors ``(c0,c2,c4)`` and ``(c1,c3,c5)``."""
    six = np.asarray(six, dtype=np.float64)
    a1, a2 = six[..., [0, 2, 4]], six[..., [1, 3, 5]]
    b1 = a1 / np.linalg.norm(a1, axis=-1, keepdims=True)
    u2 = a2 - np.sum(b1 * a2, axis=-1, keepdims=True) * b1
    b2 = u2 / np.linalg.norm(u2, axis=-1, keepdims=True)
    b3 = np.cross(b1, b2, axis=-1)
    return np.stack([b1[..., 0], b2[..., 0], b3[..., 0],
                     b1[..., 1], b2[..., 1], b3[..., 1],
                     b1[..., 2], b2[..., 2], b3[..., 2]], axis=-1)


def rot3x3_to_quat(m):
    """ts:521-545 - transforms3d route: eigenvector of the largest eigenvalue of the 4x4 K matrix, w >= 0."""
    qxx, qyx, qzx, qxy, qyy, qzy, qxz, qyz, qzz = np.asarray(m, dtype=np.float64).flat
    k = np.array([
        [qxx - qyy - qzz, 0, 0, 0],
        [qyx + qxy, qyy - qxx - qzz, 0, 0],
        [qzx + qxz, qzy + qyz, qzz - qxx - qyy, 0],
        [qyz - qzy, qzx - qxz, qxy - qyx, qxx + qyy + qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(k)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


def rot9_to_quat(m9):
    """ts:575-584 - one ``eigh`` per row, in a Python loop, as the reference does."""
    m9 = np.asarray(m9, dtype=np.float64)
    if m9.ndim == 1:
        return rot3x3_to_quat(m9.reshape(3, 3))
    out = np.zeros((m9.shape[0], 4))
    for i, r in enumerate(m9):
        out[i] = rot3x3_to_quat(r.reshape(3, 3))
    return out


def six_to_quat(six):
    """ts:471-473."""
    return rot9_to_quat(six_to_rot9(six))


def hips_sin_cos_to_quat(s, c):
    """ts:177-179 - yaw-only quaternion from ``atan2(sin, cos)``."""
    y = np.arctan2(s, c)
    z = np.zeros_like(y)
    return euler_to_quat(np.stack([z, y, z], axis=-1))


def average_quats(quats):
    """ts:32-51 - weight 1/S, samples flipped onto the hemisphere of row 0, sum, normalise."""
    quats = np.asarray(quats, dtype=np.float64)
    w = 1 / len(quats)
    q0 = quats[0]
    acc = q0 * w
    for qi in quats[1:]:
        acc = acc + (qi * -w if np.dot(qi, q0) < 0.0 else qi * w)
    return acc / np.linalg.norm(acc)
