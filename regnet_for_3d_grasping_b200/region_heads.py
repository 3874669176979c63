"""Region / refine heads of REGNet: mirrors of
multi_model/utils/pointnet2.py:123-197 (PointNet2TwoStage) and :199-254 (PointNet2Refine) with the same constructor
arguments, attribute names (=> state-dict keys of Appendix B) and outputs.

The reference feeds these heads a materialised (M, 256, G) gather of per-point features and max-pools it inside the
module (MaxPool1d).  `forward` keeps that contract; `forward_pooled` takes the already pooled (M, 256) features that
regnet_for_3d_grasping_b200.region.gather_max produces in one pass over all_feature (no (M, G, 256) tensor at all).
The per-centre MLPs are a few thousand rows of tiny 1x1 convolutions; they stay torch modules (autograd, BN batch
statistics in training) -- the work of this stage is the cropping, not these GEMMs.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _conv1x1(conv, x2d):
    """A Conv1d(kernel 1) applied to (M, C_in) rows as an fp32 GEMM.  cuDNN convolutions default to TF32 on this
    hardware (torch.backends.cudnn.allow_tf32), which costs ~1e-3 of relative accuracy; torch.matmul stays fp32."""
    return torch.addmm(conv.bias, x2d, conv.weight.view(conv.out_channels, conv.in_channels).t())


class PointNet2TwoStage(nn.Module):
    def __init__(self, num_points, input_chann, k_cls, k_reg, k_reg_theta, add_channel_flag=False):
        super().__init__()
        self.num_points = num_points
        self.k_reg = k_reg
        self.k_cls = k_cls
        self.k_reg_no_anchor = self.k_reg // self.k_cls
        self.k_reg_theta = k_reg_theta
        self.conv = nn.Conv1d(256 * (3 if add_channel_flag else 1), 1024, 1)
        self.bn = nn.BatchNorm1d(1024)
        self.conv_cls2 = nn.Conv1d(1024, 256, 1)
        self.conv_cls3 = nn.Conv1d(256, 128, 1)
        self.linear_cls = nn.Linear(128, self.k_cls)          # present (and unused) in the reference: keeps the keys
        self.conv_cls4 = nn.Conv1d(128, self.k_cls, 1)
        self.bn_cls2 = nn.BatchNorm1d(256)
        self.bn_cls3 = nn.BatchNorm1d(128)
        self.bn_cls4 = nn.BatchNorm1d(self.k_cls)
        self.conv_reg2 = nn.Conv1d(1024, 256, 1)
        self.conv_reg3 = nn.Conv1d(256, 128, 1)
        self.conv_reg4 = nn.Conv1d(128, self.k_reg, 1)
        self.bn_reg2 = nn.BatchNorm1d(256)
        self.bn_reg3 = nn.BatchNorm1d(128)
        self.bn_reg4 = nn.BatchNorm1d(self.k_reg)
        self.mp1 = nn.MaxPool1d(num_points)
        self.ap = nn.AdaptiveAvgPool1d(1)
        self.sigmod = nn.Sigmoid()

    def _heads(self, mp_x):
        m = mp_x.reshape(mp_x.shape[0], -1)                        # (M, C, 1) -> (M, C): every layer is a 1x1 convolution
        x = F.relu(self.bn(_conv1x1(self.conv, m)))
        c = F.relu(self.bn_cls2(_conv1x1(self.conv_cls2, x)))
        c = F.relu(self.bn_cls3(_conv1x1(self.conv_cls3, c)))
        x_cls = self.bn_cls4(_conv1x1(self.conv_cls4, c))
        r = F.relu(self.bn_reg2(_conv1x1(self.conv_reg2, x)))
        r = F.relu(self.bn_reg3(_conv1x1(self.conv_reg3, r)))
        r = self.bn_reg4(_conv1x1(self.conv_reg4, r))
        x_reg = r.view(r.size(0), -1, self.k_reg_no_anchor)
        x_reg[:, :, 7:] = self.sigmod(x_reg[:, :, 7:])      # scores in (0,1); in place like the reference (:189)
        return x_cls, x_reg

    def forward(self, xyz, feature):
        """xyz (M, 256, G) gathered per-point features [, feature (M, C')] -> x_cls (M,k_cls), x_reg (M,k_cls,k_reg/k_cls),
        mp_x (M, 256[+C'], 1)."""
        mp_x = self.mp1(xyz)
        if feature is not None:
            mp_x = torch.cat((mp_x, feature.view(feature.shape[0], feature.shape[1], 1)), dim=1)
        x_cls, x_reg = self._heads(mp_x)
        return x_cls, x_reg, mp_x

    def forward_pooled(self, pooled):
        """pooled (M, 256) = max over each centre's group (region.gather_max) -> same outputs as forward()."""
        mp_x = pooled.unsqueeze(-1)
        x_cls, x_reg = self._heads(mp_x)
        return x_cls, x_reg, mp_x


class PointNet2Refine(nn.Module):
    def __init__(self, num_points=2500, input_chann=3, k_cls=2, k_reg=8):
        super().__init__()
        self.num_points = num_points
        self.k_reg = k_reg
        self.k_cls = k_cls
        self.conv_formal = nn.Conv1d(384, 1024, 1)
        self.bn_formal = nn.BatchNorm1d(1024)
        self.conv_formal_cls2 = nn.Conv1d(1024, 128, 1)
        self.conv_formal_cls3 = nn.Conv1d(128, self.k_cls, 1)
        self.bn_formal_cls2 = nn.BatchNorm1d(128)
        self.bn_formal_cls3 = nn.BatchNorm1d(self.k_cls)
        self.conv_formal_reg2 = nn.Conv1d(1024, 128, 1)
        self.conv_formal_reg3 = nn.Conv1d(128, self.k_reg, 1)
        self.bn_formal_reg2 = nn.BatchNorm1d(128)
        self.bn_formal_reg3 = nn.BatchNorm1d(self.k_reg)
        self.mp1 = nn.MaxPool1d(num_points)
        self.ap = nn.AdaptiveAvgPool1d(1)
        self.sigmoid = nn.Sigmoid()

    def _heads(self, x):
        x = F.relu(self.bn_formal(_conv1x1(self.conv_formal, x.reshape(x.shape[0], -1))))
        c = F.relu(self.bn_formal_cls2(_conv1x1(self.conv_formal_cls2, x)))
        c = self.bn_formal_cls3(_conv1x1(self.conv_formal_cls3, c))
        r = F.relu(self.bn_formal_reg2(_conv1x1(self.conv_formal_reg2, x)))
        r = self.bn_formal_reg3(_conv1x1(self.conv_formal_reg3, r))
        return c, r

    def forward(self, gripper_feature, group_feature):
        """gripper_feature (M', 256, 64) [, group_feature (M', 128)] -> x_cls (M',k_cls), x_reg (M',k_reg)."""
        x = self.mp1(gripper_feature)
        if group_feature is not None:
            x = torch.cat((x, group_feature.view(group_feature.shape[0], group_feature.shape[1], 1)), dim=1)
        return self._heads(x)

    def forward_pooled(self, pooled, group_feature):
        """pooled (M', 256) from region.gather_max over the 64 closing-box points; group_feature (M', 128)."""
        x = pooled.unsqueeze(-1)
        if group_feature is not None:
            x = torch.cat((x, group_feature.view(group_feature.shape[0], group_feature.shape[1], 1)), dim=1)
        return self._heads(x)
