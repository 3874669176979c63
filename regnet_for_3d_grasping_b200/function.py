"""Operator API mirror of the reference's pn2_utils/function.py (the names REGNet's modules call).

Same callables, argument order and gradient behaviour: index-producing ops are non-differentiable
(function.py:46-48,76-78,131-133), group_points / feature_interpolate back-propagate through the scatter-add
kernels (function.py:103-107,167-172)."""
import torch

from . import pn2_ext


def gather_points(points, index):
    """function.py:11-26: points (B,C,N), index (B,M) -> (B,C,M)."""
    b, c, _ = points.shape
    return points.gather(2, index.unsqueeze(1).expand(b, c, index.size(1)))


class _FarthestPointSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, num_centroids):
        return pn2_ext.farthest_point_sample(points, num_centroids)

    @staticmethod
    def backward(ctx, *grads):
        return None, None


class _BallQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, centroids, radius, num_neighbours):
        index, count = pn2_ext.ball_query(points, centroids, radius, num_neighbours)
        return index, count

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None, None


class _GroupPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, index):
        ctx.save_for_backward(index)
        ctx.num_points = points.size(2)
        return pn2_ext.group_points_forward(points, index)

    @staticmethod
    def backward(ctx, *grads):
        (index,) = ctx.saved_tensors
        return pn2_ext.group_points_backward(grads[0], index, ctx.num_points), None


class _SearchNNDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query_xyz, key_xyz, num_neighbors):
        index, distance = pn2_ext.point_search(query_xyz, key_xyz, num_neighbors)
        return index, distance

    @staticmethod
    def backward(ctx, *grads):
        return None, None, None


class _FeatureInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, index, weight):
        ctx.save_for_backward(index, weight)
        ctx.num_inst = feature.size(2)
        return pn2_ext.interpolate_forward(feature, index, weight)

    @staticmethod
    def backward(ctx, *grads):
        index, weight = ctx.saved_tensors
        return pn2_ext.interpolate_backward(grads[0], index, weight, ctx.num_inst), None, None


farthest_point_sample = _FarthestPointSample.apply
ball_query = _BallQuery.apply
group_points = _GroupPoints.apply
search_nn_distance = _SearchNNDistance.apply
feature_interpolate = _FeatureInterpolate.apply
