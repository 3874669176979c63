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
