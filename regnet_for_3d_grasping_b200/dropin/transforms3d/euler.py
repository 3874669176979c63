"""transforms3d.euler: euler2quat for the default 'sxyz' convention (static frame, rotations about x, then y, then z),
the only one the reference uses (utils.py:436)."""
import math

import numpy as np


def euler2quat(ai, aj, ak, axes="sxyz"):
    if axes != "sxyz":
        raise NotImplementedError("only the default 'sxyz' axes convention is provided")
    ai, aj, ak = ai / 2.0, aj / 2.0, ak / 2.0
    ci, si = math.cos(ai), math.sin(ai)
    cj, sj = math.cos(aj), math.sin(aj)
    ck, sk = math.cos(ak), math.sin(ak)
    cc, cs, sc, ss = ci * ck, ci * sk, si * ck, si * sk
    return np.array([cj * cc + sj * ss, cj * sc - sj * cs, cj * ss + sj * cc, cj * cs - sj * sc])
