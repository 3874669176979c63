"""reference: pn2_utils/functions/gather_knn.py"""
import torch

import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200 import dgcnn_ext


class _GatherKNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, index):
        ctx.save_for_backward(index)
        return dgcnn_ext.gather_knn_forward(feature, index)

    @staticmethod
    def backward(ctx, grad_output):
        (index,) = ctx.saved_tensors
        return dgcnn_ext.gather_knn_backward(grad_output.contiguous(), index), None


gather_knn = _GatherKNN.apply
