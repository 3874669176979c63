import numpy as np


def Vector3dVector(values):
    """An (N, 3) float64 array (open3d's container converts to exactly that with np.asarray)."""
    a = np.array(values, dtype=np.float64)
    return a.reshape(-1, 3)


def Vector3iVector(values):
    return np.array(values, dtype=np.int32).reshape(-1, 3)


def IntVector(values=()):
    return list(int(v) for v in values)


def DoubleVector(values=()):
    return list(float(v) for v in values)
