"""PointNet2Seg -- the ScoreNet body (multi_model/utils/pointnet2.py:12-121), same constructor, attribute names
and state-dict keys.

forward():
  * eval mode  -> one call into the native plan (csrc/scorenet.cu): FPS, ball query, grouping, tcgen05 MLPs with
                  fused BN/ReLU/max-pool, 3-NN interpolation, seg head and score, all on device;
  * train mode -> the op-by-op module path (modules.py) so BatchNorm sees batch statistics, dropout is applied
                  and autograd records the graph; the point operators underneath are still this package's kernels.
Both return what the reference returns: (sparse_feature (B,256,N), x_score (B,N)); in eval mode sparse_feature is
a transposed view of the plan's point-major (B,N,256) buffer, so ScoreNetwork's `.transpose(2,1)` hands the
caller a contiguous (B,N,256) tensor.
"""
import os

import torch
import torch.nn as nn

from . import _lib, conv_train, train_ops
from .modules import PointNetSAModule, PointnetFPModule
from .nn_layers import SharedMLP
from .scorenet import NUM_CENTROIDS, NUM_NEIGHBOURS, RADIUS, ScoreNetPlan

SA_CHANNELS = ((128, 128, 256), (256, 256, 512), (512, 512, 1024))
FP_CHANNELS = ((1024, 1024), (512, 512), (256, 256, 256))
SEG_CHANNELS = (512, 256, 256, 128)


class PointNet2Seg(nn.Module):
    _SA_MODULE = PointNetSAModule
    _FP_MODULE = PointnetFPModule
    _MAX_PLANS = 4      # native plans kept per module (one per (batch, points, device)); least recently used is freed

    def __init__(self, input_chann=3, k_score=1, k_obj=2, add_channel_flag=False, dropout_prob=0.5):
        super().__init__()
        self.k_score = k_score
        feat = input_chann - 3
        skip = [feat]
        self.sa_modules = nn.ModuleList()
        for i, widths in enumerate(SA_CHANNELS):
            self.sa_modules.append(self._SA_MODULE(in_channels=feat, mlp_channels=widths, num_centroids=NUM_CENTROIDS[i],
                                                   radius=RADIUS[i], num_neighbours=NUM_NEIGHBOURS[i], use_xyz=True))
            feat = widths[-1]
            skip.append(feat)
        self.fp_modules = nn.ModuleList()
        for i, widths in enumerate(FP_CHANNELS):
            self.fp_modules.append(self._FP_MODULE(in_channels=feat + skip[-2 - i], mlp_channels=widths, num_neighbors=3))
            feat = widths[-1]
        self.mlp = SharedMLP(feat * (3 if add_channel_flag else 1), SEG_CHANNELS, ndim=1, dropout_prob=dropout_prob)
        self.conv_score = nn.Conv1d(SEG_CHANNELS[-1], self.k_score, 1)
        self.bn_score = nn.BatchNorm1d(self.k_score)
        self.sigmoid = nn.Sigmoid()
        self._fusable = (input_chann == 6 and k_score == 1 and not add_channel_flag)
        self._plans = {}
        self._geom_plans = {}
        self.engine = None  # None = library default (tcgen05); tests may force _lib.ENGINE_SIMT

    # -- op-by-op path (training) ---------------------------------------------------------------------------
    def _forward_modules(self, points, add_channel1=None, add_channel2=None):
        B, _, N = points.size()
        xyz, feature = points[:, :3, :], points[:, 3:6, :]
        geom = self._train_geometry(points)
        level_xyz, level_feature = [xyz], [feature]
        for i, sa in enumerate(self.sa_modules):
            if geom is None:
                xyz, feature = sa(xyz, feature)
            else:
                xyz, feature = sa(xyz, feature, geometry=(geom["new_xyz"][i], geom["nbr"][i]))
            level_xyz.append(xyz)
            level_feature.append(feature)
        sparse_xyz, sparse_feature = xyz, feature
        for i, fp in enumerate(self.fp_modules):
            dense_xyz, dense_feature = level_xyz[-2 - i], level_feature[-2 - i]
            if geom is None:
                sparse_feature = fp(dense_xyz, sparse_xyz, dense_feature, sparse_feature)
            else:
                sparse_feature = fp(dense_xyz, sparse_xyz, dense_feature, sparse_feature,
                                    search=(geom["nn_idx"][i], geom["nn_w"][i]))
            sparse_xyz = dense_xyz
        if add_channel1 is not None and add_channel2 is not None:
            c = sparse_feature.shape[1]
            extra = [a.view(B, 1, N).repeat(1, c, 1).float() for a in (add_channel1, add_channel2)]
            sparse_feature = torch.cat([sparse_feature] + extra, dim=1)
        x = self.mlp(sparse_feature)
        if self.training and conv_train.conv1x1_supported(self.conv_score, x):
            s = conv_train.conv1x1_train(self.conv_score, x)      # 128 -> k_score on the tcgen05 engine (no cuDNN)
            s = (train_ops.bn_relu_train(s, self.bn_score, False) if train_ops.bn_supported(s, self.bn_score)
                 else self.bn_score(s))
            x_score = s.transpose(2, 1).contiguous()
        else:
            x_score = self.bn_score(self.conv_score(x)).transpose(2, 1).contiguous()
        return sparse_feature, self.sigmoid(x_score).view(B, N)

    def _geometry_plan(self, B, N, device):
        key = (B, N, str(device))
        plan = self._geom_plans.get(key)
        if plan is None:
            while len(self._geom_plans) >= self._MAX_PLANS:
                self._geom_plans.pop(next(iter(self._geom_plans))).close()
            plan = ScoreNetPlan(B, N, device)       # no weights bound: only its geometry chain is used
            plan.set_option("defer_prefetch", 0)    # nothing to hide behind in train mode: a prefetch starts at once
            self._geom_plans[key] = plan
        return plan

    def _train_geometry(self, points):
        """Train mode on CUDA: FPS / ball query / 3-NN of every level from the native plan's geometry chain -- the same
        kernels as the operator path (bit-identical indices), but the three levels and the 3-NN searches run on the plan's
        side streams while the first MLPs are already computing, and a prefetch() issued during the previous step's
        backward takes the whole chain (FPS included) off the critical path.  None -> the modules search for themselves."""
        if not (self.training and self._fusable and points.is_cuda and points.dtype == torch.float32
                and points.size(2) >= NUM_CENTROIDS[0] and os.environ.get("REGNET_TRAIN_PLAN_GEOMETRY", "1") != "0"):
            return None
        pc = points.permute(0, 2, 1)
        if not pc.is_contiguous():
            pc = pc.contiguous()
        with torch.no_grad():
            return self._geometry_plan(pc.size(0), pc.size(1), pc.device).geometry(pc)

    # -- fused path (eval) ---------------------------------------------------------------------------------------
    _STATE_FIELDS = ("weight", "bias", "running_mean", "running_var")

    def _state_tensors(self):
        """{state-dict name: tensor} of every conv / BN tensor, read through getattr on the sub-modules.  Unlike
        state_dict() / parameters() this also works on the replicas nn.DataParallel creates (the reference's multi-GPU
        mode, utils.py:129-133): a replica's `_parameters` is empty, its weights are plain tensor attributes."""
        out = {}
        for name, mod in self.named_modules():
            if not isinstance(mod, (nn.modules.conv._ConvNd, nn.modules.batchnorm._BatchNorm)):
                continue
            for field in self._STATE_FIELDS:
                t = getattr(mod, field, None)
                if torch.is_tensor(t):
                    out[f"{name}.{field}"] = t
        return out

    def _plan_for(self, B, N, device):
        key = (B, N, str(device), self.engine)
        plan = self._plans.get(key)
        if plan is None:
            while len(self._plans) >= self._MAX_PLANS:      # a plan owns GBs of workspace: keep only the most recent shapes
                self._plans.pop(next(iter(self._plans))).close()
            plan = ScoreNetPlan(B, N, device, engine=self.engine)
            self._plans[key] = plan
        else:
            self._plans[key] = self._plans.pop(key)         # most recently used last
        tensors = self._state_tensors()
        state_key = tuple((t.data_ptr(), t._version) for t in tensors.values())
        if plan._bound_key != state_key:       # parameters changed (or first use): fold BN again and upload
            plan.bind_state({"x." + k: v for k, v in tensors.items()}, root="x.", key=state_key)
        return plan

    def _forward_fused(self, points):
        B, _, N = points.size()
        pc = points.permute(0, 2, 1)  # (B,N,6); contiguous again if it came from ScoreNetwork's permute
        plan = self._plan_for(B, N, points.device)
        all_feature, score = plan.forward(pc.float())
        return all_feature.transpose(1, 2), score

    def prefetch(self, pc):
        """Throughput mode: start the geometry chain for a future batch `pc` (B,N,6) now, overlapped with
        the forward issued next.  Later call forward on the same tensor."""
        if not self._fusable:
            return
        if self.training:
            if os.environ.get("REGNET_TRAIN_PLAN_GEOMETRY", "1") != "0" and pc.is_cuda and pc.size(1) >= NUM_CENTROIDS[0]:
                self._geometry_plan(pc.size(0), pc.size(1), pc.device).prefetch(pc)
            return
        plan = self._plan_for(pc.size(0), pc.size(1), pc.device)
        plan.prefetch(pc)

    def join_prefetch(self):
        """Make the current stream wait for the side-stream work of every outstanding prefetch()."""
        for plan in list(self._plans.values()) + list(self._geom_plans.values()):
            plan.join_prefetch()

    def forward(self, points, add_channel1=None, add_channel2=None):
        fused = (not self.training and self._fusable and add_channel1 is None and points.is_cuda
                 and points.size(2) >= NUM_CENTROIDS[0])
        if fused:
            return self._forward_fused(points)
        return self._forward_modules(points, add_channel1, add_channel2)

    def __getstate__(self):  # plans hold native handles: never pickle them (torch.save(model), train.py:175-178)
        state = self.__dict__.copy()
        state["_plans"] = {}
        state["_geom_plans"] = {}
        return state
