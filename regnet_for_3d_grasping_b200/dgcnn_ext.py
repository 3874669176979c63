"""Drop-in for the reference's `dgcnn_ext` (multi_model/utils/pn2_utils/functions/csrc/main.cpp:3-6).

gather_knn is group_points with N' == N (functions/csrc/gather_knn_kernel.cu:27-50, 100-153); it is only
imported, never executed, by REGNet (pn2_utils/modules.py:6 -> dead EdgeFeatureInterpolator)."""
from . import pn2_ext as _ops


def gather_knn_forward(input, index):
    return _ops.group_points_forward(input, index)


def gather_knn_backward(grad_output, index):
    # the reference sizes grad_input by grad_output.size(2) (gather_knn_kernel.cu:106,120)
    return _ops.group_points_backward(grad_output, index, grad_output.size(2))
