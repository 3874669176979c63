"""CPU restatement of the deterministic parts of the reference's region stage (torch CPU, same op order).

TEST INFRASTRUCTURE ONLY.  Follows /root/reference/dataset_utils/get_regiondataset.py:
  positives / centre selection   :354-434  (the FPS branch is exact; the two random branches are checked by property)
  ball membership                :279-295  (sqrt of separately rounded squares, non-strict <=)
and /root/reference/multi_model/gripper_region_network.py:532-544 for the masked sampler's thresholds.
Parity is "mask-level exact + distributional": the reference draws from numpy's global RNG seeded with the wall
clock (train.py:59), so its random picks are not reproducible by anyone.
"""
import torch

from oracle import pn2_oracle


def select_score_center_fps_branch(pc, pre_score, center_num, score_thre):
    """Per cloud: (count of positives, centre indices or None when count <= center_num)."""
    out = []
    for b in range(pc.shape[0]):
        mask = pre_score[b] > score_thre
        pos = torch.nonzero(mask).view(-1)
        if len(pos) > center_num:
            cur = pc[b, pos, :3]
            idx = pn2_oracle.farthest_point_sample(cur.view(1, -1, 3).transpose(2, 1), center_num).view(-1)
            out.append((len(pos), pos[idx]))
        else:
            out.append((len(pos), None))
    return out


def ball_mask(all_points, center, radius):
    """(center_num, N) bool: get_regiondataset.py:287-294 written with broadcasting instead of repeats."""
    d = all_points[None, :, :3] - center[:, None, :3]
    dist = torch.sqrt(torch.mul(d[..., 0], d[..., 0]) + torch.mul(d[..., 1], d[..., 1]) + torch.mul(d[..., 2], d[..., 2]))
    return dist <= radius


# ---- GraspRegionNet + RefineNet, inference call (multi_model/gripper_region_network.py, ground_grasp=None) ------------
# Functional restatement from a state dict; pinned by tests/golden/ref_py_region_net.npz, which the REAL reference module
# produced on CPU (oracle/gen_golden_cpu.py: `.cuda()` made a no-op, np.random.choice replaced by a fixed rule).
import math

import torch.nn.functional as F


def _cb(sd, x, conv, bn, relu):
    """Conv1d(k=1, with bias) + BatchNorm1d (eval) [+ ReLU]  (utils/pointnet2.py:161-189, 223-251)."""
    y = F.conv1d(x, sd[conv + ".weight"], sd[conv + ".bias"])
    y = F.batch_norm(y, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"], False, 0.0, 1e-5)
    return F.relu(y) if relu else y


def two_stage_head(sd, p, mp_x):
    """PointNet2TwoStage after the max-pool (utils/pointnet2.py:165-197): mp_x (M,256,1) -> x_cls (M,4), x_reg (M,4,10)."""
    x = _cb(sd, mp_x, p + "conv", p + "bn", True)
    c = _cb(sd, x, p + "conv_cls2", p + "bn_cls2", True)
    c = _cb(sd, c, p + "conv_cls3", p + "bn_cls3", True)
    c = _cb(sd, c, p + "conv_cls4", p + "bn_cls4", False)
    r = _cb(sd, x, p + "conv_reg2", p + "bn_reg2", True)
    r = _cb(sd, r, p + "conv_reg3", p + "bn_reg3", True)
    r = _cb(sd, r, p + "conv_reg4", p + "bn_reg4", False)
    x_reg = r.view(r.shape[0], -1, 10).clone()
    x_reg[:, :, 7:] = torch.sigmoid(x_reg[:, :, 7:])
    return c.view(c.shape[0], c.shape[1]), x_reg


def refine_head(sd, p, x):
    """PointNet2Refine after max-pool + concat (utils/pointnet2.py:227-254): x (M,384,1) -> cls (M,2), reg (M,10)."""
    x = _cb(sd, x, p + "conv_formal", p + "bn_formal", True)
    c = _cb(sd, x, p + "conv_formal_cls2", p + "bn_formal_cls2", True)
    c = _cb(sd, c, p + "conv_formal_cls3", p + "bn_formal_cls3", False)
    r = _cb(sd, x, p + "conv_formal_reg2", p + "bn_formal_reg2", True)
    r = _cb(sd, r, p + "conv_formal_reg3", p + "bn_formal_reg3", False)
    return c.view(c.shape[0], -1), r.view(r.shape[0], -1)


def anchors_for(centers):
    """gripper_region_network.py:31-44, 553-586: four fp16-rounded orientation templates, theta 0."""
    s = math.sqrt(3) / 3
    t = torch.tensor([[s, s, s, 0], [s, s, -s, 0], [s, -s, -s, 0], [s, -s, s, 0]], dtype=torch.float32).half().float()
    return torch.cat([centers.view(-1, 1, 3).expand(-1, 4, 3), t.view(1, 4, 4).expand(centers.shape[0], 4, 4)], dim=-1)


def decode_first_stage(x_reg, anchors, x_cls, radius):
    """gripper_region_network.py:69-90."""
    pick = x_cls.argmax(dim=1)
    rows = torch.arange(x_reg.shape[0])
    g, t = x_reg[rows, pick], anchors[rows, pick]
    axis = g[:, 3:6] + t[:, 3:6]
    norm = torch.sqrt(torch.sum(axis * axis, dim=1) + 1e-12).view(-1, 1)
    return torch.cat([g[:, :3] * radius + t[:, :3], axis / norm, math.pi * (g[:, 6:7] + t[:, 6:7]), g[:, 7:]], dim=-1)


def closing_box(group_points, grasp, params):
    """gripper_region_network.py:445-528: points in the gripper frame and the strict six-plane membership mask."""
    width, height, depth = params
    M = grasp.shape[0]

    def unit(v, fallback, eps=True):
        n = torch.norm(v, dim=1) + (1e-12 if eps else 0.0)
        out = v / n.view(-1, 1)
        out[n == 0] = torch.tensor(fallback)
        return out

    ay = unit(grasp[:, 3:6].float(), [0.0, 1.0, 0.0])
    ct, st = torch.cos(grasp[:, 6].float()), torch.sin(grasp[:, 6].float())
    z, o = torch.zeros(M), torch.ones(M)
    R1 = torch.stack([ct, z, -st, z, o, z, st, z, ct], dim=1).view(M, 3, 3)
    ax = unit(torch.stack([ay[:, 1], -ay[:, 0], z], dim=1), [1.0, 0.0, 0.0])
    az = unit(torch.cross(ax, ay, dim=1), [0.0, 0.0, 1.0], eps=False)
    approach = unit(torch.bmm(torch.stack([ax, ay, az], dim=2), R1)[:, :, 0], [1.0, 0.0, 0.0])
    rot = torch.stack([approach, ay, torch.cross(approach, ay, dim=1)], dim=1)
    pts = torch.bmm(rot, (group_points[:, :, :3].float() - grasp[:, None, :3].float()).permute(0, 2, 1)).permute(0, 2, 1)
    x, y, zz = pts[:, :, 0], pts[:, :, 1], pts[:, :, 2]
    mask = (x > 0) & (x < depth / 2) & (y > -width / 2) & (y < width / 2) & (zz > -height / 2) & (zz < height / 2)
    return pts, mask


def sample_rows_fixed_rule(mask, num, min_count=5):
    """The fixture's stand-in for np.random.choice (gen_golden_cpu.deterministic_choice) applied to a membership mask:
    more than `num` members -> the first `num`; more than `min_count` -> members[(7 i + 3) mod count]; else rejected (-1)."""
    out = torch.full((mask.shape[0], num), -1, dtype=torch.int64)
    for m in range(mask.shape[0]):
        members = torch.nonzero(mask[m]).view(-1)
        c = len(members)
        if c > num:
            out[m] = members[:num]
        elif c > min_count:
            out[m] = members[(torch.arange(num) * 7 + 3) % c]
    return out


def region_net_forward(sd, inp, params, group_num=16, gripper_num=8, score_thre=0.4, radius=0.06):
    """The 16-tuple's deterministic members for the inference call, from a GripperRegionNetwork state dict."""
    B, N, C = inp["all_feature"].shape
    N_C, N_GM = inp["pc_group_more_index"].shape[1:]
    flat = inp["all_feature"].reshape(B * N, C)
    gi = (inp["pc_group_index"].long() + torch.arange(B).view(B, 1, 1) * N).view(B * N_C, -1)
    mp = flat[gi].max(dim=1)[0].view(B * N_C, C, 1)                     # gather + MaxPool1d(group_num)
    x_cls, x_reg = two_stage_head(sd, "extrat_feature_region.", mp)
    next_grasp = decode_first_stage(x_reg, anchors_for(inp["center_pc"][:, :, :3].reshape(-1, 3).float()), x_cls, radius)
    pts, mask = closing_box(inp["pc_group_more"].view(B * N_C, N_GM, 6), next_grasp, params)
    pick = sample_rows_fixed_rule(mask, gripper_num)
    ok = pick[:, 0] >= 0
    gmask = torch.nonzero(ok).view(-1)
    inall = inp["pc_group_more_index"].view(B * N_C, N_GM).long().gather(1, pick.clamp(min=0))
    rows = (inall + (torch.arange(B * N_C) // N_C).view(-1, 1) * N)[gmask]
    gf = flat[rows].max(dim=1)[0]                                        # (M', 256)
    centre_half = mp.view(-1, 128)[gmask]                                # the reference's (2M,128) re-view, :343
    cls, reg = refine_head(sd, "extrat_feature_refine.", torch.cat([gf, centre_half], dim=1).unsqueeze(-1))
    stage1 = next_grasp[gmask]
    final = stage1.clone()
    final[:, :3] = final[:, :3] + reg[:, :3] * radius
    final[:, 3:] = final[:, 3:] + reg[:, 3:]
    positive = cls.argmax(dim=1) == 1
    class_select = torch.nonzero(positive).view(-1)
    score_select = torch.nonzero(positive & (final[:, 7] > score_thre)).view(-1)
    return dict(next_grasp=next_grasp, gripper_mask=gmask, gripper_pc_index=torch.where(ok[:, None], pick, torch.full_like(pick, -1)),
                gripper_pc_index_inall=torch.where(ok[:, None], inall, torch.full_like(inall, -1)),
                sel_class=final[class_select], sel_score=final[score_select], sel_stage2=stage1[class_select],
                final_mask=gmask[class_select], final_mask_sthre=gmask[score_select], closing_mask=mask)
