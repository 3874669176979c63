"""transforms3d.quaternions: quat2mat, axangle2quat (quaternions are (w, x, y, z))."""
import math

import numpy as np


def quat2mat(q):
    """Rotation matrix of a quaternion (normalised first; the zero quaternion gives the identity, like transforms3d)."""
    w, x, y, z = (float(v) for v in q)
    n = w * w + x * x + y * y + z * z
    if n < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / n
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ = w * X, w * Y, w * Z
    xX, xY, xZ = x * X, x * Y, x * Z
    yY, yZ, zZ = y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY],
                     [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                     [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def axangle2quat(vector, theta, is_normalized=False):
    """Quaternion of a rotation by `theta` about `vector`."""
    v = np.asarray(vector, dtype=np.float64)
    if not is_normalized:
        v = v / math.sqrt(float(np.dot(v, v)))
    t2 = theta / 2.0
    return np.concatenate(([math.cos(t2)], v * math.sin(t2)))
