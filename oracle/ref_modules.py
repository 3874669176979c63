"""Plain-PyTorch restatement of the reference's ScoreNet forward, driven by a state dict.

TEST INFRASTRUCTURE ONLY (see oracle/pn2_oracle.c).  Used as (a) the feature oracle for the parity tests
(fp32 or fp64 on CPU), (b) the CPU baseline timed by bench.py (`cpu_baseline`, `--impl reference`), and
(c) with oracle/_ref's kernels as `ext`, the "reference CUDA kernels on B200" side baseline.

It restates, in functional form, the reference call chain (paths relative to /root/reference/multi_model/):
  score_network.py:31-53                       ScoreNetwork.forward
  utils/pointnet2.py:86-121                    PointNet2Seg.forward (arch constants :40-46)
  utils/pn2_utils/modules.py:210-246           PointNetSAModule.forward
  utils/pn2_utils/modules.py:39-56             QueryGrouper.forward  (concat order [xyz_rel, feature])
  utils/pn2_utils/modules.py:104-131,500-509   FeatureInterpolator / PointnetFPModule ([interp, dense])
  utils/pn2_utils/nn/modules/mlp.py:95-106     SharedMLP (conv1x1 no-bias -> BN -> ReLU per layer)
  utils/pn2_utils/function.py:11-26            gather_points
gen_golden_cpu.py checks this restatement against the real reference modules imported from /root/reference.

`ext` is any object with the reference's pn2_ext entry points (oracle.pn2_oracle.as_pn2_ext() on CPU,
oracle/_ref's pn2_ext_ref on a GPU).
"""
import torch
import torch.nn.functional as F

# architecture constants, pointnet2.py:40-46
NUM_CENTROIDS = (5120, 1024, 256)
RADIUS = (0.02, 0.08, 0.32)
NUM_NEIGHBOURS = (64, 64, 64)
SA_CHANNELS = ((128, 128, 256), (256, 256, 512), (512, 512, 1024))
FP_CHANNELS = ((1024, 1024), (512, 512), (256, 256, 256))
SEG_CHANNELS = (512, 256, 256, 128)
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


from regnet_for_3d_grasping_b200.weights import random_scorenet_state, scorenet_state_shapes  # noqa: E402,F401


def gather_points(points, index):
    b, c, _ = points.shape
    return points.gather(2, index.unsqueeze(1).expand(b, c, index.size(1)))


def _shared_mlp(sd, prefix, nlayers, x, training, dtype):
    for j in range(nlayers):
        w = sd[f"{prefix}.{j}.conv.weight"].to(dtype)
        x = F.conv2d(x, w) if w.dim() == 4 else F.conv1d(x, w)
        p = f"{prefix}.{j}.bn."
        x = F.batch_norm(x, sd[p + "running_mean"].to(dtype), sd[p + "running_var"].to(dtype),
                         sd[p + "weight"].to(dtype), sd[p + "bias"].to(dtype), training, BN_MOMENTUM, BN_EPS)
        x = F.relu(x)
    return x


def scorenet_forward(sd, pc, ext, dtype=torch.float32, training=False, keep=False, ops_dtype=torch.float32,
                     num_centroids=NUM_CENTROIDS, radius=RADIUS, num_neighbours=NUM_NEIGHBOURS):
    """pc (B,N,>=6) -> (all_feature (B,N,256) [transpose view], score (B,N), intermediates dict).

    The neighbour-search ops always run in fp32 (`ops_dtype`), as in REGNet; `dtype` selects the arithmetic
    of the MLP / interpolation part (fp64 gives the high-precision feature oracle).  Dropout (train mode,
    pointnet2.py:78) is not applied: train-mode comparisons use all_feature only.  `num_centroids`/`radius`
    default to the reference's constants; the parity tests shrink them to run small clouds."""
    root = "extrat_featurePN2."
    points = pc[:, :, :6].permute(0, 2, 1)
    xyz = points[:, :3, :].to(ops_dtype)
    feature = points[:, 3:6, :].to(dtype)
    inter_xyz, inter_feature, dbg = [xyz], [feature], {}
    for i in range(3):
        index = ext.farthest_point_sample(xyz, num_centroids[i])
        new_xyz = gather_points(xyz, index)
        nbr, cnt = ext.ball_query(xyz, new_xyz, radius[i], num_neighbours[i])
        b, _, m = new_xyz.shape
        k = nbr.size(2)
        flat = nbr.reshape(b, 1, m * k)
        gxyz = xyz.gather(2, flat.expand(b, 3, m * k)).view(b, 3, m, k) - new_xyz.unsqueeze(-1)
        gfeat = feature.gather(2, flat.expand(b, feature.size(1), m * k)).view(b, feature.size(1), m, k)
        grouped = torch.cat([gxyz.to(dtype), gfeat], dim=1)
        h = _shared_mlp(sd, f"{root}sa_modules.{i}.mlp", 3, grouped, training, dtype)
        feature = h.max(dim=3)[0]
        if keep:
            dbg[f"fps{i}"], dbg[f"bq{i}"], dbg[f"bqcnt{i}"], dbg[f"sa{i}"] = index, nbr, cnt, feature
        xyz = new_xyz
        inter_xyz.append(xyz)
        inter_feature.append(feature)
    sparse_xyz, sparse_feature = xyz, feature
    for i in range(3):
        dense_xyz, dense_feature = inter_xyz[-2 - i], inter_feature[-2 - i]
        index, dist = ext.point_search(dense_xyz, sparse_xyz, 3)
        inv = 1.0 / torch.clamp(dist, min=1e-10)                      # modules.py:120 (fp32, like the reference)
        weight = inv / inv.sum(dim=2, keepdim=True)
        b, c, _ = sparse_feature.shape
        nd = index.size(1)
        if dtype == torch.float32:
            interp = ext.interpolate_forward(sparse_feature, index, weight)   # the kernel's own fma chain
        else:
            nb = sparse_feature.gather(2, index.reshape(b, 1, nd * 3).expand(b, c, nd * 3)).view(b, c, nd, 3)
            wd = weight.to(dtype)
            interp = nb[..., 0] * wd[:, None, :, 0]                   # interpolate_kernel.cu:165-170, k = 0,1,2
            interp = interp + nb[..., 1] * wd[:, None, :, 1]
            interp = interp + nb[..., 2] * wd[:, None, :, 2]
        x = torch.cat([interp, dense_feature], dim=1)
        sparse_feature = _shared_mlp(sd, f"{root}fp_modules.{i}.mlp", len(FP_CHANNELS[i]), x, training, dtype)
        if keep:
            dbg[f"nn{i}"], dbg[f"nnd{i}"], dbg[f"fp{i}"] = index, dist, sparse_feature
        sparse_xyz = dense_xyz
    x = _shared_mlp(sd, f"{root}mlp", 4, sparse_feature, training, dtype)
    s = F.conv1d(x, sd[root + "conv_score.weight"].to(dtype), sd[root + "conv_score.bias"].to(dtype))
    p = root + "bn_score."
    s = F.batch_norm(s, sd[p + "running_mean"].to(dtype), sd[p + "running_var"].to(dtype), sd[p + "weight"].to(dtype),
                     sd[p + "bias"].to(dtype), training, BN_MOMENTUM, BN_EPS)
    score = torch.sigmoid(s.transpose(2, 1).contiguous()).view(pc.size(0), -1)
    return sparse_feature.transpose(2, 1), score, dbg
