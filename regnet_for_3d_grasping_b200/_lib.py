"""ctypes binding of libregnet_b200.so (include/regnet_b200.h).  No fallback: if the library is missing or a
call fails, this raises."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libregnet_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "regnet_b200.h")

ENGINE_TC = 0
ENGINE_SIMT = 1

_lib = None

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p


class ScoreNetConfig(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("num_points", ctypes.c_int32), ("num_centroids", ctypes.c_int32 * 3),
                ("radius", ctypes.c_float * 3), ("num_neighbours", ctypes.c_int32 * 3), ("engine", ctypes.c_int32),
                ("use_side_stream", ctypes.c_int32)]


_STRIDED = [c_ptr, c_i64, c_i64, c_i64]
_SIGNATURES = {
    "regnet_last_error": (ctypes.c_char_p, []),
    "regnet_abi_version": (c_int, []),
    "regnet_device_arch": (c_int, [ctypes.POINTER(c_int)] * 3),
    "regnet_check_index_errors": (c_int, []),
    "regnet_farthest_point_sample": (c_int, _STRIDED + [c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_farthest_point_sample_ex": (c_int, _STRIDED + [c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_int, c_int, c_ptr]),
    "regnet_ball_query": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, c_f32, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "regnet_farthest_point_sample_f64": (c_int, _STRIDED + [c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_ball_query_f64": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, ctypes.c_double, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_point_search_f64": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_search_workspace_bytes": (c_i64, [c_int, c_int]),
    "regnet_ball_query_ws": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, c_f32, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_point_search_ws": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_group_points_forward": (c_int, _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_group_points_backward": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_point_search": (c_int, _STRIDED + _STRIDED + [c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_interpolate_forward": (c_int, _STRIDED + [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_interpolate_backward": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_gather_knn_forward": (c_int, _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_gather_knn_backward": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_scorenet_create": (c_int, [ctypes.POINTER(ScoreNetConfig), ctypes.POINTER(c_ptr)]),
    "regnet_scorenet_destroy": (c_int, [c_ptr]),
    "regnet_scorenet_workspace_bytes": (c_i64, [c_ptr]),
    "regnet_scorenet_set_layer": (c_int, [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "regnet_scorenet_forward": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "regnet_scorenet_prefetch": (c_int, [c_ptr, c_ptr, c_ptr]),
    "regnet_scorenet_join_prefetch": (c_int, [c_ptr, c_ptr]),
    "regnet_scorenet_geometry": (c_int, [c_ptr, c_ptr, c_ptr]),
    "regnet_scorenet_set_option": (c_int, [c_ptr, ctypes.c_char_p, c_int]),
    "regnet_scorenet_intermediate": (c_int, [c_ptr, ctypes.c_char_p, ctypes.POINTER(c_ptr), ctypes.POINTER(c_i64)]),
    "regnet_scorenet_launch_count": (c_int, [c_ptr]),
    "regnet_scorenet_set_profiling": (c_int, [c_ptr, c_int]),
    "regnet_scorenet_profile": (c_int, [c_ptr, ctypes.c_char_p, c_i64]),
    "regnet_bn_workspace_bytes": (c_i64, [c_int, c_int, c_i64]),
    "regnet_bn_relu_train_forward": (c_int, [c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_f32, c_f32, c_int, c_ptr, c_ptr,
                                             c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_bn_relu_train_backward": (c_int, [c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_ptr,
                                              c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_bn_relu_max64_train_forward": (c_int, [c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_f32, c_f32, c_int, c_ptr, c_ptr,
                                                   c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_bn_relu_max64_train_backward": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                                    c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_split_planes": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "regnet_split_weight": (c_int, [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_conv1x1_train": (c_int, [c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_int,
                                     c_ptr]),
    "regnet_conv1x1_wgrad_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_i64]),
    "regnet_conv1x1_train_wgrad": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_i64, c_ptr, c_ptr, c_i64,
                                           c_int, c_ptr]),
    "regnet_bn_finalize_moments": (c_int, [c_ptr, c_int, ctypes.c_double, c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr,
                                           c_ptr, c_ptr, c_ptr, c_ptr]),
    "regnet_bn_apply_ex": (c_int, [c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_int, c_f32, ctypes.c_uint64, c_ptr, c_ptr,
                                   c_ptr, c_ptr]),
    "regnet_bn_backward_ex": (c_int, [c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32,
                                      ctypes.c_uint64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_conv1x1_train_dgrad_bnreduce": (c_int, [c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_int, c_int, c_ptr,
                                                    c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32, ctypes.c_uint64, c_ptr,
                                                    c_int, c_ptr]),
    "regnet_bn_backward_from_sums": (c_int, [c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_f32,
                                             ctypes.c_uint64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_sa_group_planes": (c_int, _STRIDED + [c_ptr] + _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                                            c_ptr]),
    "regnet_fp_interp_planes": (c_int, _STRIDED + _STRIDED + [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr,
                                                              c_ptr]),
    "regnet_group_points_backward_strided": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr,
                                                     c_ptr]),
    "regnet_interpolate_backward_strided": (c_int, [c_ptr, c_i64, c_int, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_ptr,
                                                    c_ptr]),
    "regnet_sa_gather_linear": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_sa_scatter_linear": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_fp_gather_linear": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_ptr, c_int, c_int, c_int, c_int,
                                        c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_fp_dense_wgrad": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_sa0_input_moments": (c_int, _STRIDED + [c_ptr] + _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_int, c_ptr,
                                                              c_ptr, c_ptr]),
    "regnet_sa0_backward_finalize": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_int, ctypes.c_double, c_ptr, c_ptr, c_ptr, c_ptr]),
    "regnet_sa0_apply_planes": (c_int, _STRIDED + [c_ptr] + _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                                             c_int, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_sa0_backward_sums": (c_int, _STRIDED + [c_ptr] + _STRIDED + [c_ptr, c_int, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr,
                                                              c_ptr, c_int, c_int, c_ptr, c_ptr]),
    "regnet_bn_apply_max64": (c_int, [c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_int, c_ptr, c_ptr, c_ptr]),
    "regnet_bn_max64_backward_ex": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_int,
                                            c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_maxpool64_forward": (c_int, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "regnet_maxpool64_backward": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr]),
    "regnet_select_score_center_workspace": (c_i64, [c_int, c_int, c_int]),
    "regnet_select_score_center": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_f32, ctypes.c_uint64, c_ptr, c_ptr, c_ptr,
                                           c_ptr, c_i64, c_ptr]),
    "regnet_ball_crop_sample": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_f32, c_int, ctypes.c_uint64, c_ptr, c_ptr,
                                        c_ptr, c_ptr]),
    "regnet_ball_crop_workspace_bytes": (c_i64, [c_int, c_int]),
    "regnet_ball_crop_sample_ws": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_f32, c_int, ctypes.c_uint64, c_ptr, c_ptr,
                                           c_ptr, c_ptr, c_i64, c_ptr]),
    "regnet_closing_box_mask": (c_int, [c_ptr, c_int, c_int, c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_f32, c_ptr,
                                        c_ptr]),
    "regnet_mask_sample": (c_int, [c_ptr, c_int, c_int, c_int, c_int, ctypes.c_uint64, c_ptr, c_ptr, c_ptr]),
    "regnet_gather_max": (c_int, [c_ptr, c_ptr, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_mlp_layer": (c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_int, c_int, c_int, c_int, c_int, c_ptr, c_ptr]),
    "regnet_linear_planes": (c_int, [c_ptr, c_ptr, c_int, c_i64, c_int, c_ptr, c_ptr, c_int, c_int, c_ptr, c_ptr, c_int, c_ptr,
                                     c_int, c_ptr, c_ptr, c_int, c_ptr]),
    "regnet_sa0_chain": (c_int, [c_ptr, c_ptr, c_ptr, c_int, c_int, c_int] + [c_ptr] * 9 + [c_ptr, c_ptr, c_int, c_ptr]),
}


def declared_symbols():
    """Every function the public header declares (used by the symbol-export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(regnet_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built. Run `python -m regnet_for_3d_grasping_b200.build` (needs nvcc); "
            "there is no CPU or PyTorch fallback for these operators.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().regnet_last_error()
        raise RuntimeError((msg or b"unknown error").decode("utf-8", "replace"))


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
