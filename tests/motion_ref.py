"""numpy restatement of the reference's key-frame interpolation (Transform.h:129-240,417-436; Quaternion.h:136-188), shared by the
GPU motion-blur tests and the CPU cross-check against the oracle's C++ restatement."""
import numpy as np


def quat_from_axis_angle(axis, angle):
    if angle == 0.0:
        return np.array([0, 0, 0, 1.0])
    axis = np.asarray(axis, np.float64) / np.linalg.norm(axis)
    return np.concatenate([axis * np.sin(angle / 2), [np.cos(angle / 2)]])


def quat_mix(q0, q1, t):
    q0, q1 = q0 / np.linalg.norm(q0), q1 / np.linalg.norm(q1)
    d = float(q0 @ q1)
    if d < 0:
        q1, d = -q1, -d
    if d > 0.9995:
        q = q0 + (q1 - q0) * t
        return q / np.linalg.norm(q)
    th0 = np.arccos(d)
    th = th0 * t
    return q0 * (np.cos(th) - d * np.sin(th) / np.sin(th0)) + q1 * (np.sin(th) / np.sin(th0))


def quat_matrix(q):
    x, y, z, w = q
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]])


def pose_matrix(keys, time):
    """keys: list of (axis, angle, scale[3], translation[3], time). 3x4 object->world matrix at `time` (interpolateFrames)."""
    def parts(k):
        return quat_from_axis_angle(k[0], k[1]), np.asarray(k[2], np.float64), np.asarray(k[3], np.float64)
    if len(keys) == 1 or time == 0.0 or time <= keys[0][4]:
        q, s, tr = parts(keys[0])
    elif time >= keys[-1][4]:
        q, s, tr = parts(keys[-1])
    else:
        i = next(i for i in range(1, len(keys)) if keys[i - 1][4] <= time <= keys[i][4])
        u = (time - keys[i - 1][4]) / (keys[i][4] - keys[i - 1][4])
        (q0, s0, t0), (q1, s1, t1) = parts(keys[i - 1]), parts(keys[i])
        q, s, tr = quat_mix(q0, q1, u), s0 * (1 - u) + s1 * u, t0 * (1 - u) + t1 * u
    m = np.zeros((3, 4))
    m[:, :3] = quat_matrix(q) * s[None, :]
    m[:, 3] = tr
    return m
