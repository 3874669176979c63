"""ctypes front end of oracle/pn2_oracle.c, shaped like the reference's `pn2_ext` module
(multi_model/utils/pn2_utils/csrc/main.cpp:6-14) but running on CPU torch tensors.

TEST INFRASTRUCTURE ONLY -- see the header of pn2_oracle.c.  The product package never imports this.

`as_pn2_ext()` returns an object with the 7 reference entry points so the *unmodified* reference Python
modules (modules.py, pointnet2.py, score_network.py) can be driven on CPU for golden-vector generation
(oracle/gen_golden_cpu.py) and the restated modules in oracle/ref_modules.py can be timed as the CPU baseline.
"""
import ctypes
import os
import subprocess
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pn2_oracle.c")
LIB = os.path.join(HERE, "_build", "libpn2_oracle.so")

_lib = None


def build(force=False):
    if os.path.exists(LIB) and not force and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -ffp-contract=off: the fma placement is explicit in the source (see header of pn2_oracle.c)
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c99", "-ffp-contract=off", "-fopenmp", SRC, "-o", LIB, "-lm"]
    subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        i64, f32p, i64p = ctypes.c_int64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int64)
        _lib.oracle_fps.argtypes = [f32p, i64, i64, i64, i64p]
        _lib.oracle_ball_query.argtypes = [f32p, f32p, i64, i64, i64, ctypes.c_float, i64, i64p, i64p]
        _lib.oracle_three_nn.argtypes = [f32p, f32p, i64, i64, i64, i64p, f32p]
        _lib.oracle_interpolate.argtypes = [f32p, i64p, f32p, i64, i64, i64, i64, f32p]
        _lib.oracle_interpolate_bw.argtypes = [f32p, i64p, f32p, i64, i64, i64, i64, f32p]
        _lib.oracle_group.argtypes = [f32p, i64p, i64, i64, i64, i64, i64, f32p]
        _lib.oracle_group_bw.argtypes = [f32p, i64p, i64, i64, i64, i64, i64, f32p]
        _lib.oracle_fps_block_size.argtypes = [i64]
        for n in ("oracle_fps", "oracle_ball_query", "oracle_three_nn", "oracle_interpolate",
                  "oracle_interpolate_bw", "oracle_group", "oracle_group_bw", "oracle_fps_block_size",
                  "oracle_num_threads"):
            getattr(_lib, n).restype = ctypes.c_int
    return _lib


def _f32(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _i64(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int64))


def _aos(points):
    """(B,3,N) any stride -> (B,N,3) contiguous fp32, as the reference host code does."""
    if points.dim() != 3 or points.size(1) != 3:
        raise RuntimeError("expected a (B, 3, N) tensor")
    return points.detach().to(torch.float32).transpose(1, 2).contiguous()


def num_threads():
    return lib().oracle_num_threads()


def set_threads(n):
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline should use every host core."""
    lib().oracle_set_threads.argtypes = [ctypes.c_int]
    lib().oracle_set_threads.restype = None
    lib().oracle_set_threads(int(n))


def farthest_point_sample(points, num_centroids):
    p = _aos(points)
    B, N, _ = p.shape
    if num_centroids <= 0 or N < num_centroids:
        raise RuntimeError("farthest_point_sample: need 0 < num_centroids <= num_points")
    out = torch.zeros(B, num_centroids, dtype=torch.int64)
    rc = lib().oracle_fps(_f32(p), B, N, num_centroids, _i64(out))
    if rc != 0:
        raise RuntimeError(f"oracle_fps failed rc={rc}")
    return out


def ball_query(points, centroids, radius, num_neighbours):
    p, c = _aos(points), _aos(centroids)
    B, N, _ = p.shape
    M = c.shape[1]
    index = torch.zeros(B, M, num_neighbours, dtype=torch.int64)
    count = torch.zeros(B, M, dtype=torch.int64)
    lib().oracle_ball_query(_f32(p), _f32(c), B, N, M, ctypes.c_float(radius), num_neighbours, _i64(index), _i64(count))
    return [index, count]


def point_search(query_xyz, key_xyz, num_neighbours):
    if num_neighbours != 3:
        raise RuntimeError("point_search only supports 3 neighbours")
    q, k = _aos(query_xyz), _aos(key_xyz)
    B, Nq, _ = q.shape
    Nk = k.shape[1]
    if Nk < 3:
        raise RuntimeError("point_search needs at least 3 keys")
    index = torch.zeros(B, Nq, 3, dtype=torch.int64)
    dist = torch.zeros(B, Nq, 3, dtype=torch.float32)
    lib().oracle_three_nn(_f32(q), _f32(k), B, Nq, Nk, _i64(index), _f32(dist))
    return [index, dist]


def interpolate_forward(input, index, weight):
    x = input.detach().to(torch.float32).contiguous()
    idx = index.contiguous()
    w = weight.detach().to(torch.float32).contiguous()
    B, C, Ns = x.shape
    Nd = idx.shape[1]
    out = torch.zeros(B, C, Nd, dtype=torch.float32)
    rc = lib().oracle_interpolate(_f32(x), _i64(idx), _f32(w), B, C, Ns, Nd, _f32(out))
    if rc != 0:
        raise RuntimeError("interpolate_forward: index out of range")
    return out


def interpolate_backward(grad_output, index, weight, num_inst):
    g = grad_output.detach().to(torch.float32).contiguous()
    idx = index.contiguous()
    w = weight.detach().to(torch.float32).contiguous()
    B, C, Nd = g.shape
    out = torch.zeros(B, C, num_inst, dtype=torch.float32)
    lib().oracle_interpolate_bw(_f32(g), _i64(idx), _f32(w), B, C, num_inst, Nd, _f32(out))
    return out


def group_points_forward(input, index):
    x = input.detach().to(torch.float32).contiguous()
    idx = index.contiguous()
    B, C, N = x.shape
    _, M, K = idx.shape
    out = torch.zeros(B, C, M, K, dtype=torch.float32)
    rc = lib().oracle_group(_f32(x), _i64(idx), B, C, N, M, K, _f32(out))
    if rc != 0:
        raise RuntimeError("group_points_forward: index out of range")
    return out


def group_points_backward(grad_output, index, num_points):
    g = grad_output.detach().to(torch.float32).contiguous()
    idx = index.contiguous()
    B, C, M, K = g.shape
    out = torch.zeros(B, C, num_points, dtype=torch.float32)
    lib().oracle_group_bw(_f32(g), _i64(idx), B, C, num_points, M, K, _f32(out))
    return out


def as_pn2_ext():
    m = types.ModuleType("pn2_ext")
    for f in (farthest_point_sample, ball_query, point_search, interpolate_forward, interpolate_backward,
              group_points_forward, group_points_backward):
        setattr(m, f.__name__, f)
    return m


def as_dgcnn_ext():
    """functions/csrc/gather_knn_kernel.cu:27-50,100-153 -- same maths as group_points."""
    m = types.ModuleType("dgcnn_ext")
    m.gather_knn_forward = group_points_forward
    m.gather_knn_backward = lambda g, idx: group_points_backward(g, idx, idx.shape[1])
    return m


# ---- independent numpy statements of the same semantics (slow; used to cross-check the C file) ----------

def np_fps_tierule(points_aos, M):
    """FPS via the closed-form tie rule of SURVEY.md Appendix A.1 (bit-reversed slot), no block emulation."""
    p = np.asarray(points_aos, dtype=np.float32)
    N = p.shape[0]
    block = lib().oracle_fps_block_size(N)
    nbits = block.bit_length() - 1
    j = np.arange(N)
    t = j % block
    rev = np.zeros(N, dtype=np.int64)
    for b in range(nbits):
        rev |= ((t >> b) & 1) << (nbits - 1 - b)
    prio = rev * (N // block + 2) + j // block  # smaller wins
    temp = np.full(N, np.inf, dtype=np.float32)
    out = np.zeros(M, dtype=np.int64)
    cur = 0
    for i in range(1, M):
        d = p - p[cur]
        dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
        t0 = (dy * dy).astype(np.float32)
        # fma emulated in float64 (exact product, one rounding): products of fp32 are exact in fp64
        t1 = (dx.astype(np.float64) * dx.astype(np.float64) + t0.astype(np.float64)).astype(np.float32)
        t2 = (dz.astype(np.float64) * dz.astype(np.float64) + t1.astype(np.float64)).astype(np.float32)
        temp = np.minimum(temp, t2)
        mx = temp.max()
        if mx > 0:
            cand = np.nonzero(temp == mx)[0]
            cur = int(cand[np.argmin(prio[cand])])
        out[i] = cur
    return out
