"""Per-module error of the train-mode forward (chained MLPs on the tcgen05 engine) against the same modules evaluated in
float64 by torch on the SAME inputs.  GPU box only."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.score_network import ScoreNetwork  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
net = ScoreNetwork(training=True).cuda()
net.load_state_dict(weights.random_scorenet_state(seed=3))
net.train()
pn = net.extrat_featurePN2
pc = torch.from_numpy(synth.batch("table", [1, 2], 6144)).cuda()
points = pc.permute(0, 2, 1)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()


def both(mlp, x, pooled):
    os.environ["REGNET_TRAIN_TORCH"] = "0"
    m1 = copy.deepcopy(mlp)
    y1 = m1.forward_max_over_neighbours(x) if pooled else m1(x)
    os.environ["REGNET_TRAIN_TORCH"] = "1"
    m2 = copy.deepcopy(mlp).double()
    y2 = m2(x.double())
    if pooled:
        y2 = y2.max(dim=3)[0]
    m3 = copy.deepcopy(mlp)
    y3 = m3(x)
    if pooled:
        y3 = y3.max(dim=3)[0]
    os.environ["REGNET_TRAIN_TORCH"] = "0"
    # per-layer check of the chain: statistics
    for i, (b1, b2) in enumerate(zip(m1, m2)):
        print(f"      layer {i}: running_mean err {rel(b1.bn.running_mean, b2.bn.running_mean):.2e} running_var err {rel(b1.bn.running_var, b2.bn.running_var):.2e}")
    return y1, y2, y3


with torch.no_grad():
    from regnet_for_3d_grasping_b200 import function as F_
    xyz, feature = points[:, :3, :], points[:, 3:6, :]
    level_xyz, level_feature = [xyz], [feature]
    for i, sa in enumerate(pn.sa_modules):
        new_xyz = F_.gather_points(xyz, sa.sampler(xyz))
        group_feature, _ = sa.grouper(new_xyz, xyz, feature, use_xyz=True)
        y1, y2, y3 = both(sa.mlp, group_feature, True)
        print(f"sa{i}: in {tuple(group_feature.shape)} chain err {rel(y1, y2):.3e} torch-fp32 err {rel(y3, y2):.3e}", flush=True)
        xyz, feature = new_xyz, y2.float()
        level_xyz.append(xyz)
        level_feature.append(feature)
    sparse_xyz, sparse_feature = xyz, feature
    for i, fp in enumerate(pn.fp_modules):
        dense_xyz, dense_feature = level_xyz[-2 - i], level_feature[-2 - i]
        x = fp.interpolator(dense_xyz, sparse_xyz, dense_feature, sparse_feature)
        y1, y2, y3 = both(fp.mlp, x, False)
        print(f"fp{i}: in {tuple(x.shape)} chain err {rel(y1, y2):.3e} torch-fp32 err {rel(y3, y2):.3e}", flush=True)
        sparse_feature, sparse_xyz = y2.float(), dense_xyz
