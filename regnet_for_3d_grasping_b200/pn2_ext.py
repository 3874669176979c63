"""Drop-in for the reference's pybind module `pn2_ext` (multi_model/utils/pn2_utils/csrc/main.cpp:6-14).

Same seven names, positional signatures, shapes, dtypes and error behaviour (RuntimeError on bad arguments), but
every call goes through the C ABI of libregnet_b200.so (include/regnet_b200.h) on torch's current CUDA stream.
Put regnet_for_3d_grasping_b200/dropin on PYTHONPATH and the reference's `import pn2_ext`
(multi_model/utils/pn2_utils/function.py:2) resolves to this module -- see INTEGRATION.md.

The north-star spellings (furthest_point_sample, three_nn, three_interpolate) are exported as aliases.
There is no CPU path, exactly like the reference (CHECK_CUDA everywhere).
"""
import ctypes

import torch

from . import _lib


def _need_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")


def _is_f64(*tensors):
    """The reference dispatches over float and double (AT_DISPATCH_FLOATING_TYPES).  float64 inputs take the generic
    double-precision kernels (csrc/generic_ops.cu) for the search operators and exact torch gathers / scatter-adds for
    the value operators -- complete, not tuned: REGNet itself only passes float32."""
    if all(isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 for t in tensors):
        return True
    return False


def _need_index(t, name):
    if not t.is_cuda or t.dtype != torch.int64:
        raise RuntimeError(f"{name} must be a CUDA int64 tensor")


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _strided3(t):
    return (_p(t), t.stride(0), t.stride(1), t.stride(2))


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def farthest_point_sample(points, num_centroids):
    """csrc/sampling_kernel.cu:126-170.  points (B,3,N) any stride -> index (B,M) int64."""
    f64 = _is_f64(points)
    if not f64:
        _need_cuda_f32(points, "points")
    if points.dim() != 3 or points.size(1) != 3:
        raise RuntimeError("points.size(1) does not equal to 3")
    B, _, N = points.shape
    M = int(num_centroids)
    if M <= 0:
        raise RuntimeError("num_centroids is not greater than 0")
    if N < M:
        raise RuntimeError("num_points is not greater or equal than num_centroids")
    index = torch.empty(B, M, dtype=torch.int64, device=points.device)
    if B == 0:
        return index
    with torch.cuda.device(points.device):
        if f64:
            _lib.check(_lib.load().regnet_farthest_point_sample_f64(*_strided3(points), B, N, M, _p(index), _stream()))
        else:
            _lib.check(_lib.load().regnet_farthest_point_sample(*_strided3(points), B, N, M, _p(index), None, _stream()))
    return index


def ball_query(points, centroids, radius, num_neighbours):
    """csrc/ball_query_kernel.cu:87-131.  -> [index (B,M,K) int64, count (B,M) int64]."""
    f64 = _is_f64(points, centroids)
    if not f64:
        _need_cuda_f32(points, "points")
        _need_cuda_f32(centroids, "centroids")
    if points.dim() != 3 or points.size(1) != 3:
        raise RuntimeError("points.size(1) does not equal to 3")
    if centroids.dim() != 3 or centroids.size(1) != 3:
        raise RuntimeError("centroids.size(1) does not equal to 3")
    B, _, N = points.shape
    M = centroids.size(2)
    K = int(num_neighbours)
    index = torch.empty(B, M, K, dtype=torch.int64, device=points.device)
    count = torch.empty(B, M, dtype=torch.int64, device=points.device)
    if B == 0 or M == 0:
        return [index, count]
    with torch.cuda.device(points.device):
        lib = _lib.load()
        if f64:
            _lib.check(lib.regnet_ball_query_f64(*_strided3(points), *_strided3(centroids), B, N, M, float(radius), K,
                                                 _p(index), _p(count), _stream()))
        elif K == 64 and 4096 <= N <= 65536:
            # big clouds: the uniform-grid kernels (same results); scratch comes from torch's caching allocator
            nbytes = int(lib.regnet_search_workspace_bytes(B, N))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
            _lib.check(lib.regnet_ball_query_ws(*_strided3(points), *_strided3(centroids), B, N, M, float(radius), K,
                                                _p(index), _p(count), _p(ws), nbytes, _stream()))
        else:
            _lib.check(lib.regnet_ball_query(*_strided3(points), *_strided3(centroids), B, N, M, float(radius), K,
                                             _p(index), _p(count), None, _stream()))
    return [index, count]


def group_points_forward(input, index):
    """csrc/grouping_kernel.cu:29-51.  input (B,C,N), index (B,M,K) -> (B,C,M,K)."""
    if _is_f64(input):
        B, C, _ = input.shape
        _, M, K = index.shape
        return input.gather(2, index.reshape(B, 1, M * K).expand(B, C, M * K)).view(B, C, M, K)
    _need_cuda_f32(input, "input")
    _need_index(index, "index")
    if input.dim() != 3 or index.dim() != 3 or index.size(0) != input.size(0):
        raise RuntimeError("group_points_forward: expected input (B,C,N) and index (B,M,K)")
    B, C, N = input.shape
    _, M, K = index.shape
    index = index.contiguous()
    out = torch.empty(B, C, M, K, dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(input.device):
        _lib.check(_lib.load().regnet_group_points_forward(*_strided3(input), _p(index), B, C, N, M, K, _p(out), _stream()))
    return out


def group_points_backward(grad_output, index, num_points):
    """csrc/grouping_kernel.cu:103-149.  grad (B,C,M,K) -> (B,C,N)."""
    if _is_f64(grad_output):
        B, C, M, K = grad_output.shape
        out = torch.zeros(B, C, int(num_points), dtype=torch.float64, device=grad_output.device)
        return out.scatter_add_(2, index.reshape(B, 1, M * K).expand(B, C, M * K), grad_output.reshape(B, C, M * K))
    _need_cuda_f32(grad_output, "grad_output")
    _need_index(index, "index")
    if grad_output.dim() != 4 or index.dim() != 3 or tuple(index.shape) != (grad_output.size(0), grad_output.size(2), grad_output.size(3)):
        raise RuntimeError("group_points_backward: expected grad_output (B,C,M,K) and index (B,M,K)")
    B, C, M, K = grad_output.shape
    g = grad_output.contiguous()
    index = index.contiguous()
    out = torch.empty(B, C, int(num_points), dtype=torch.float32, device=g.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().regnet_group_points_backward(_p(g), _p(index), B, C, int(num_points), M, K, _p(out), _stream()))
    return out


def point_search(query_xyz, key_xyz, num_neighbours):
    """csrc/interpolate_kernel.cu:88-128.  -> [index (B,Nq,3) int64, squared distance (B,Nq,3)]."""
    f64 = _is_f64(query_xyz, key_xyz)
    if not f64:
        _need_cuda_f32(query_xyz, "query_xyz")
        _need_cuda_f32(key_xyz, "key_xyz")
    if key_xyz.size(0) != query_xyz.size(0) or query_xyz.size(1) != 3 or key_xyz.size(1) != 3:
        raise RuntimeError("point_search: expected (B,3,N1) and (B,3,N2)")
    if int(num_neighbours) != 3:
        raise RuntimeError("num_neighbours does not equal to K")
    B, _, Nq = query_xyz.shape
    Nk = key_xyz.size(2)
    if Nk < 3:
        raise RuntimeError("num_key is not greater or equal than num_neighbours")
    index = torch.empty(B, Nq, 3, dtype=torch.int64, device=query_xyz.device)
    dist = torch.empty(B, Nq, 3, dtype=query_xyz.dtype, device=query_xyz.device)
    if B == 0 or Nq == 0:
        return [index, dist]
    with torch.cuda.device(query_xyz.device):
        lib = _lib.load()
        if f64:
            _lib.check(lib.regnet_point_search_f64(*_strided3(query_xyz), *_strided3(key_xyz), B, Nq, Nk, 3, _p(index),
                                                   _p(dist), _stream()))
        elif 4096 <= Nk <= 65536:
            nbytes = int(lib.regnet_search_workspace_bytes(B, Nk))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=query_xyz.device)
            _lib.check(lib.regnet_point_search_ws(*_strided3(query_xyz), *_strided3(key_xyz), B, Nq, Nk, 3, _p(index),
                                                  _p(dist), _p(ws), nbytes, _stream()))
        else:
            _lib.check(lib.regnet_point_search(*_strided3(query_xyz), *_strided3(key_xyz), B, Nq, Nk, 3, _p(index),
                                               _p(dist), _stream()))
    return [index, dist]


def interpolate_forward(input, index, weight):
    """csrc/interpolate_kernel.cu:187-232.  input (B,C,Ns), index/weight (B,Nd,3) -> (B,C,Nd)."""
    if _is_f64(input, weight):
        B, C, _ = input.shape
        Nd = index.size(1)
        g = input.gather(2, index.reshape(B, 1, Nd * 3).expand(B, C, Nd * 3)).view(B, C, Nd, 3)
        w = weight.unsqueeze(1)
        return (g[..., 0] * w[..., 0] + g[..., 1] * w[..., 1]) + g[..., 2] * w[..., 2]
    _need_cuda_f32(input, "input")
    _need_index(index, "index")
    _need_cuda_f32(weight, "weight")
    B, C, Ns = input.shape
    if index.size(0) != B or index.size(2) != 3 or tuple(weight.shape) != tuple(index.shape):
        raise RuntimeError("interpolate_forward: expected index and weight of shape (B,Nd,3)")
    Nd = index.size(1)
    index = index.contiguous()
    weight = weight.contiguous()
    out = torch.empty(B, C, Nd, dtype=torch.float32, device=input.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(input.device):
        _lib.check(_lib.load().regnet_interpolate_forward(*_strided3(input), _p(index), _p(weight), B, C, Ns, Nd, _p(out), _stream()))
    return out


def interpolate_backward(grad_output, index, weight, num_inst):
    """csrc/interpolate_kernel.cu:292-337.  grad (B,C,Nd) -> (B,C,Ns)."""
    if _is_f64(grad_output, weight):
        B, C, Nd = grad_output.shape
        out = torch.zeros(B, C, int(num_inst), dtype=torch.float64, device=grad_output.device)
        contrib = (grad_output.unsqueeze(-1) * weight.unsqueeze(1)).reshape(B, C, Nd * 3)
        return out.scatter_add_(2, index.reshape(B, 1, Nd * 3).expand(B, C, Nd * 3), contrib)
    _need_cuda_f32(grad_output, "grad_output")
    _need_index(index, "index")
    _need_cuda_f32(weight, "weight")
    B, C, Nd = grad_output.shape
    if index.size(0) != B or index.size(2) != 3 or weight.size(0) != B or weight.size(1) != Nd or weight.size(2) != 3:
        raise RuntimeError("interpolate_backward: expected index and weight of shape (B,Nd,3)")
    g = grad_output.contiguous()
    index = index.contiguous()
    weight = weight.contiguous()
    out = torch.empty(B, C, int(num_inst), dtype=torch.float32, device=g.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().regnet_interpolate_backward(_p(g), _p(index), _p(weight), B, C, int(num_inst), Nd, _p(out), _stream()))
    return out


def check_index_errors():
    """Synchronising check for out-of-range indices seen by the gather/scatter ops since the last call
    (the reference device-asserts instead)."""
    _lib.check(_lib.load().regnet_check_index_errors())


# north-star / erikwijmans spellings
furthest_point_sample = farthest_point_sample
three_nn = point_search
three_interpolate = interpolate_forward
