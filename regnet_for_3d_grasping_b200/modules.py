"""Host-side mirror of the PointNet++ modules REGNet instantiates (multi_model/utils/pn2_utils/modules.py):
FarthestPointSampler :11-29, QueryGrouper :32-59, FeatureInterpolator :98-134, PointNetSAModule :176-252,
PointnetFPModule :480-512.  Same constructor arguments, attribute names (=> same state-dict keys) and
outputs; the operators underneath are this package's CUDA kernels (function.py -> pn2_ext -> C ABI).

This is the op-by-op path (used in training, where BN needs batch statistics and autograd needs the graph).
The eval-mode fast path does not go through these forwards: see scorenet.py."""
import torch
from torch import nn

from . import conv_train
from . import function as _F
from .nn_layers import SharedMLP


class FarthestPointSampler(nn.Module):
    def __init__(self, num_centroids):
        super().__init__()
        self.num_centroids = num_centroids

    def forward(self, points):
        with torch.no_grad():
            return _F.farthest_point_sample(points, self.num_centroids)

    def extra_repr(self):
        return 'num_centroids={:d}'.format(self.num_centroids)


class QueryGrouper(nn.Module):
    def __init__(self, radius, num_neighbours):
        super().__init__()
        assert radius > 0.0 and num_neighbours > 0
        self.radius = radius
        self.num_neighbours = num_neighbours

    def forward(self, new_xyz, xyz, feature, use_xyz, index=None):
        if index is None:
            with torch.no_grad():
                index, _ = _F.ball_query(xyz, new_xyz, self.radius, self.num_neighbours)
        group_xyz = _F.group_points(xyz, index) - new_xyz.unsqueeze(-1)   # centre on the centroid
        if feature is None:
            return group_xyz, group_xyz
        group_feature = _F.group_points(feature, index)
        if use_xyz:
            group_feature = torch.cat([group_xyz, group_feature], dim=1)  # xyz channels first
        return group_feature, group_xyz

    def extra_repr(self):
        return 'radius={}, num_neighbours={}'.format(self.radius, self.num_neighbours)


class FeatureInterpolator(nn.Module):
    def __init__(self, num_neighbors, eps=1e-10):
        super().__init__()
        self.num_neighbors = num_neighbors
        self._eps = eps

    def search(self, dense_xyz, sparse_xyz):
        """3-NN indices and normalised inverse-distance weights (modules.py:117-122)."""
        with torch.no_grad():
            index, distance = _F.search_nn_distance(dense_xyz, sparse_xyz, self.num_neighbors)
            inv = 1.0 / torch.clamp(distance, min=self._eps)
            weight = inv / torch.sum(inv, dim=2, keepdim=True)
        return index, weight

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature, search=None):
        index, weight = search if search is not None else self.search(dense_xyz, sparse_xyz)
        out = _F.feature_interpolate(sparse_feature, index, weight)
        if dense_feature is not None:
            out = torch.cat([out, dense_feature], dim=1)                  # interpolated channels first
        return out

    def extra_repr(self):
        return 'num_neighbours={:d}, eps={}'.format(self.num_neighbors, self._eps)


class PointNetSAModule(nn.Module):
    """Set abstraction: FPS -> ball query -> group -> shared MLP -> max over neighbours."""

    def __init__(self, in_channels, mlp_channels, num_centroids, radius, num_neighbours, use_xyz):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.num_centroids = num_centroids
        self.use_xyz = use_xyz
        self.mlp = SharedMLP(in_channels + (3 if use_xyz else 0), mlp_channels, ndim=2, bn=True)
        self.sampler = FarthestPointSampler(num_centroids) if num_centroids > 0 else None
        if num_neighbours < 0:
            assert radius < 0.0
            self.grouper = None
        else:
            assert num_neighbours > 0 and radius > 0.0
            self.grouper = QueryGrouper(radius, num_neighbours)

    def forward(self, xyz, feature=None, geometry=None):
        """geometry (optional, not in the reference): (new_xyz (B,3,M), ball-query index (B,M,K) int64) computed elsewhere
        for exactly this `xyz` (PointNet2Seg passes the native plan's results in train mode)."""
        if self.num_centroids == 0:       # one group holding every point, centred on the origin
            assert self.grouper is None
            new_xyz = xyz.new_zeros(xyz.size(0), 3, 1)
            group_feature = feature.unsqueeze(2)
            if self.use_xyz:
                group_feature = torch.cat([xyz.unsqueeze(2), group_feature], dim=1)
        else:
            if geometry is not None:
                new_xyz, index = geometry
            else:
                new_xyz = xyz if self.num_centroids == -1 else _F.gather_points(xyz, self.sampler(xyz))
                index = None
            if self.training and self.use_xyz and feature is not None and xyz.is_cuda:
                # train mode on CUDA: grouping + centring + concat written straight as the MLP's operand planes, the MLP
                # and the max over the neighbours as one autograd function (conv_train.py)
                if index is None:
                    with torch.no_grad():
                        index, _ = _F.ball_query(xyz, new_xyz, self.grouper.radius, self.grouper.num_neighbours)
                if conv_train.sa_grouped_supported(self.mlp, xyz, new_xyz, feature, index):
                    return new_xyz, conv_train.sa_grouped_chain_train(self.mlp, xyz, new_xyz, feature, index)
            group_feature, _ = self.grouper(new_xyz, xyz, feature, use_xyz=self.use_xyz, index=index)
        new_feature = self.mlp.forward_max_over_neighbours(group_feature)   # torch.max(self.mlp(x), 3)[0], modules.py:245
        return new_xyz, new_feature

    def init_weights(self, init_fn=None):
        self.mlp.init_weights(init_fn)

    def extra_repr(self):
        return 'num_centroids={:d}, use_xyz={}'.format(self.num_centroids, self.use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation -> concat skip features -> shared MLP."""

    def __init__(self, in_channels, mlp_channels, num_neighbors):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.mlp = SharedMLP(in_channels, mlp_channels, ndim=1, bn=True)
        if num_neighbors == 0:
            self.interpolator = None
        elif num_neighbors == 3:
            self.interpolator = FeatureInterpolator(num_neighbors)
        else:
            raise ValueError('Expected value 1 or 3, but {} given.'.format(num_neighbors))

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature, search=None):
        """search (optional, not in the reference): (index, weight) of the 3-NN interpolation computed elsewhere."""
        if self.interpolator is None:
            assert sparse_xyz.size(2) == 1 and sparse_feature.size(2) == 1
            x = torch.cat([sparse_feature.expand(-1, -1, dense_xyz.size(2)), dense_feature], dim=1)
        else:
            if self.training and sparse_feature.is_cuda:
                index, weight = search if search is not None else self.interpolator.search(dense_xyz, sparse_xyz)
                if conv_train.fp_interp_supported(self.mlp, sparse_feature, dense_feature, index, weight):
                    return conv_train.fp_interp_chain_train(self.mlp, sparse_feature, dense_feature, index, weight)
                search = (index, weight)
            x = self.interpolator(dense_xyz, sparse_xyz, dense_feature, sparse_feature, search=search)
        return self.mlp(x)

    def init_weights(self, init_fn=None):
        self.mlp.init_weights(init_fn)
