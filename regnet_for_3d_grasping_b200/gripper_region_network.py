"""GraspRegionNet + RefineNet of REGNet on device: mirror of
multi_model/gripper_region_network.py with the same constructor, `forward` signature, 16-tuple result and state-dict keys
(Appendix B), for the inference call of test.py:137-140 (`ground_grasp=None`) and the training call (train.py:240-243).

What changes underneath (nothing a caller can see except speed):
  * R3  the (B*N_C, N_G, 256) feature gather + MaxPool1d (:385-395, pointnet2.py:161) is one `gather_max` kernel; the
        materialised gather (1 GB at the test configuration) never exists;
  * R5  the anchor decode (:69-90) is vectorised -- the reference walks `final_mask` with a Python loop of per-element
        device writes (:79-80);
  * R6  `get_gripper_region_transform` (:436-550) keeps its batched frame algebra but replaces the per-grasp Python loop
        (nonzero + len() synchronisation + host RNG + H2D copy per grasp, :532-544) by one masked-sampler kernel;
  * R7  the refine stage gathers + max-pools its 64 closing-box points per grasp with the same kernel and keeps the
        reference's `view(-1, 128)` quirk on the pooled centre features (:343).
Random draws (which 64 of the points inside the closing box) come from the device generator of region.py; the reference
uses numpy's wall-clock-seeded global RNG, so only the distribution can match (tests/test_gpu_region.py).

The training branches (`ground_grasp` given: anchor classification / regression losses, :92-199, :217-309; the "loss
bookkeeping" row of SURVEY.md section 8(f)) are vectorised restatements pinned against the reference's own methods
(tests/golden/ref_py_region_losses.npz); the class-balanced subsets keep the reference's host draws (np.random.choice).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import region
from .region_heads import PointNet2Refine, PointNet2TwoStage


def _enumerate_templates():
    """The four orientation anchors (+-sqrt(3)/3 ...), theta = 0, rounded through fp16 like the reference (:553-586):
    (1, 4, 1, 4) half."""
    s = math.sqrt(3) / 3
    t_r = torch.tensor([[s, s, s], [s, s, -s], [s, -s, -s], [s, -s, s]], dtype=torch.float32).view(1, 4, 1, 3)
    t_theta = torch.zeros(1, 4, 1, 1, dtype=torch.float32)
    return torch.cat([t_r, t_theta], dim=3).half()


def compute_cos_sim(a, b):
    """1 - cos(a, b) per row (:588-604): rows (N,3) -> (N,1)."""
    eps = 1e-12
    ab = (a * b).sum(dim=1)
    a2 = (a * a).sum(dim=1) + eps
    b2 = (b * b).sum(dim=1) + eps
    return (1.0 - ab / torch.sqrt(a2 * b2)).view(-1, 1)


def _unit_rows(v, fallback, add_eps=True):
    """v / (|v| [+ 1e-12]) per row; rows whose norm is exactly 0 become `fallback` (the reference's fix-ups :462-476)."""
    n = torch.norm(v, dim=1)
    if add_eps:
        n = n + 1e-12
    out = v / n.view(-1, 1)
    zero = n == 0
    if zero.any():
        out[zero] = torch.tensor(fallback, dtype=v.dtype, device=v.device)
    return out


def closing_box_frame(grasp):
    """Rows of the world -> gripper rotation for grasps (M, >=7) = (centre, axis_y, theta, ...): (M,3,3) whose rows are
    approach, axis_y, minor normal (:445-503)."""
    M = grasp.shape[0]
    axis_y = _unit_rows(grasp[:, 3:6].float(), [0.0, 1.0, 0.0])
    angle = grasp[:, 6].float()
    cos_t, sin_t = torch.cos(angle), torch.sin(angle)
    zero, one = torch.zeros_like(cos_t), torch.ones_like(cos_t)
    R1 = torch.stack([cos_t, zero, -sin_t, zero, one, zero, sin_t, zero, cos_t], dim=1).view(M, 3, 3)
    axis_x = _unit_rows(torch.stack([axis_y[:, 1], -axis_y[:, 0], zero], dim=1), [1.0, 0.0, 0.0])
    axis_z = _unit_rows(torch.cross(axis_x, axis_y, dim=1), [0.0, 0.0, 1.0], add_eps=False)
    frame = torch.bmm(torch.stack([axis_x, axis_y, axis_z], dim=2), R1)
    approach = _unit_rows(frame[:, :, 0], [1.0, 0.0, 0.0])
    minor = torch.cross(approach, axis_y, dim=1)
    return torch.stack([approach, axis_y, minor], dim=1)


def closing_box_points(group_points, grasp, gripper_params):
    """(points in the gripper frame (M,G,3), membership mask (M,G)): the six strict half-space tests 0 < x < depth/2,
    |y| < width/2, |z| < height/2 (:505-528)."""
    widths, height, depths = gripper_params
    rot = closing_box_frame(grasp)
    rel = group_points[:, :, :3].float() - grasp[:, None, 0:3].float()
    pcs_t = torch.bmm(rot, rel.permute(0, 2, 1)).permute(0, 2, 1)
    x_limit = depths.float().view(-1, 1) / 2 if isinstance(depths, torch.Tensor) else depths / 2
    y_limit = widths.float().view(-1, 1) / 2 if isinstance(widths, torch.Tensor) else widths / 2
    z_limit = height / 2
    x, y, z = pcs_t[:, :, 0], pcs_t[:, :, 1], pcs_t[:, :, 2]
    mask = (x > 0) & (x < x_limit) & (y > -y_limit) & (y < y_limit) & (z > -z_limit) & (z < z_limit)
    return pcs_t, mask


def get_gripper_region_transform(group_points, group_index, grasp, region_num, gripper_params, seed=None, sampler=None):
    """Points of each grasp's closing area, in the gripper frame (reference :436-550).

    group_points (M,G,6), group_index (M,G) indices into the cloud, grasp (M,>=7) -> gripper_pc (M,region_num,6),
    gripper_pc_index (M,region_num) positions inside the group, gripper_pc_index_inall (M,region_num) cloud indices,
    true_mask_index (M',) = grasps with more than 5 points in the box.  Rejected rows hold -1.
    `sampler(mask) -> (M, region_num) int64` replaces the device sampler (tests use a deterministic rule)."""
    M, G, C = group_points.shape
    widths, height, depths = gripper_params
    rot = closing_box_frame(grasp)
    centre = grasp[:, 0:3].float()
    if group_points.is_cuda and group_points.dtype == torch.float32 and M > 0:
        # one pass over the crops: nothing but the membership mask is written (the reference's bmm materialises the
        # transformed (M,G,3) crop, 98 MB at the test configuration, and makes six comparison passes over it)
        half = lambda v: (v.float().reshape(-1) / 2) if isinstance(v, torch.Tensor) else v / 2
        mask = region.closing_box_mask(group_points, centre, rot, half(depths), half(widths), height / 2)
    else:
        _, mask = closing_box_points(group_points, grasp, gripper_params)
    if sampler is not None:
        index = sampler(mask.bool())
    else:
        index = region.sample_mask_rows(mask, region_num, min_count=5, seed=seed)
    ok = index[:, 0] >= 0
    safe = index.clamp(min=0)
    # the reference allocates these with torch.full(..., -1), i.e. as int64, so the float crop it writes into gripper_pc
    # is truncated (nothing consumes gripper_pc) -- same dtype and values here.  Only the picked points
    # are transformed into the gripper frame.
    chosen = group_points.float().gather(1, safe[:, :, None].expand(M, region_num, C))
    chosen_t = torch.bmm(rot, (chosen[:, :, :3] - centre[:, None, :]).permute(0, 2, 1)).permute(0, 2, 1)
    picked = torch.cat([chosen_t, chosen[:, :, 3:]], dim=-1)
    gripper_pc = torch.where(ok[:, None, None], picked.long(), torch.full_like(picked, -1).long())
    gripper_pc_index = torch.where(ok[:, None], index, torch.full_like(index, -1))
    inall = group_index.long().gather(1, safe)
    gripper_pc_index_inall = torch.where(ok[:, None], inall, torch.full_like(inall, -1))
    true_mask_index = torch.nonzero(ok).view(-1)
    return gripper_pc, gripper_pc_index, gripper_pc_index_inall, true_mask_index


def _gather_max(all_feature, index):
    """max over each group's feature rows, (B,N,C) x (B,N_C,G) -> (B,N_C,C): the one-pass kernel, or -- when a gradient
    has to flow back into all_feature (joint training of ScoreNet and the region heads) -- the reference's own
    gather + max expression (:385-395), which autograd understands."""
    if torch.is_grad_enabled() and all_feature.requires_grad:
        B, N, C = all_feature.shape
        add = torch.arange(B, device=index.device).view(B, 1, 1) * N
        rows = all_feature.reshape(-1, C)[(index.long() + add).reshape(-1)]
        return rows.view(B, index.shape[1], index.shape[2], C).max(dim=2)[0]
    return region.gather_max(all_feature, index.long())


class GripperRegionNetwork(nn.Module):
    def __init__(self, training, group_num, gripper_num, grasp_score_threshold, radius, reg_channel):
        super().__init__()
        self.group_number = group_num
        self.templates = _enumerate_templates()          # plain attribute, not a buffer -- like the reference (:14)
        self.anchor_number = self.templates.shape[1] * self.templates.shape[2]
        self.gripper_number = gripper_num
        self.grasp_score_thre = grasp_score_threshold
        self.is_training_refine = training
        self.radius = radius
        self.reg_channel = reg_channel
        self.extrat_feature_region = PointNet2TwoStage(num_points=group_num, input_chann=6, k_cls=self.anchor_number,
                                                       k_reg=self.reg_channel * self.anchor_number,
                                                       k_reg_theta=self.anchor_number)
        self.extrat_feature_refine = PointNet2Refine(num_points=gripper_num, input_chann=6, k_cls=2, k_reg=self.reg_channel)
        self.criterion_cos = nn.CosineEmbeddingLoss(reduction="mean")
        self.criterion_cls = nn.CrossEntropyLoss(reduction="mean")
        self.smooth_l1_loss = nn.SmoothL1Loss(reduction="mean")
        self.sample_seed = None      # int -> reproducible closing-box sampling (otherwise drawn from torch's CPU generator)
        self._sampler = None         # test hook, see get_gripper_region_transform

    # ---- R5: anchors and first-stage decode ---------------------------------------------------------------------------
    def _enumerate_anchors(self, centers):
        """centers (M,3) -> (M, anchor_number, 7) = (x, y, z, rx, ry, rz, theta) (:31-44)."""
        t = self.templates.to(centers.device).float().view(1, self.anchor_number, 4)
        return torch.cat([centers.view(-1, 1, 3).expand(-1, self.anchor_number, 3),
                          t.expand(centers.shape[0], -1, -1)], dim=-1)

    def decode_first_stage(self, x_reg, anchors, x_cls):
        """(:69-90) best anchor per centre, centre = reg[:3] * radius + anchor, axis = normalise(reg[3:6] + anchor),
        theta = pi * (reg[6] + anchor), scores = reg[7:].  x_reg (M,A,10), anchors (M,A,7), x_cls (M,A) ->
        next_grasp (M,10), chosen anchor (M,), chosen anchor rows (M,7)."""
        predict = torch.max(x_cls.transpose(1, 0), dim=0)[1]
        rows = torch.arange(x_reg.shape[0], device=x_reg.device)
        g = x_reg[rows, predict]                      # == first_grasp.transpose(1,0).view(-1,10)[predict * M + i]
        t = anchors[rows, predict]
        axis = g[:, 3:6] + t[:, 3:6]
        norm = torch.sqrt((axis * axis).sum(dim=1) + 1e-12).view(-1, 1)
        next_grasp = torch.cat([g[:, :3] * self.radius + t[:, :3], axis / norm, math.pi * (g[:, 6:7] + t[:, 6:7]),
                                g[:, 7:]], dim=-1)
        return next_grasp, predict, t

    def _cos_loss(self, a, b):
        """criterion_cos(a, b, ones): mean(1 - cos) -- the reference passes an (n, 1) ones target, which torch 1.8
        broadcast and current torch rejects; the value it meant is this."""
        return self.criterion_cos(a, b, a.new_ones(a.shape[0]))

    def compute_loss(self, first_grasp, anchors, first_cls, ground):
        """Reference method (:46-199).  first_grasp (M,A,10), anchors (M,A,7), first_cls (M,A), ground (B,N_C,10) | None ->
        next_grasp (m,10), loss_tuple, correct_tuple, next_gt (m,10) | None, tt_gt (m,7) | None, gmask (m,) with m = the
        centres that have a ground-truth grasp (all of them without ground).

        Vectorised: the reference addresses the (anchor, centre) pair through a transposed flat index built by Python
        loops of per-element device writes (:79-80, :108-109), loops over anchors with `.sum()` synchronisations and
        prints some forty tensors per call; here the pairs are addressed as [centre, anchor] directly.  The balanced
        anchor sampling keeps the reference's host draws (np.random.choice, same order, same arguments :121-127)."""
        if ground is None:
            next_grasp, _, _ = self.decode_first_stage(first_grasp, anchors, first_cls)
            gmask = torch.arange(first_grasp.shape[0], device=first_grasp.device)
            return next_grasp, (None, None), (None, None, None, None), None, None, gmask
        D = ground.shape[2]
        flat = ground.reshape(-1, D)
        gmask = torch.nonzero(flat[:, -1] != -1).view(-1)
        if gmask.numel() == 0:
            # a batch without a single labelled centre: the reference's means over empty selections are NaN and its refine
            # stage then sees no grasp (returns None); same here, without tripping over empty concatenations
            nan = first_grasp.new_tensor(float("nan"))
            loss = first_grasp.sum() * 0.0 + nan
            zero = first_grasp.new_tensor(0.0)
            return (first_grasp.new_zeros((0, D)), (loss,) + (nan,) * 9, (zero, zero), first_grasp.new_zeros((0, D)),
                    first_grasp.new_zeros((0, 7)), gmask)
        anchors, first_grasp, first_cls = anchors[gmask], first_grasp[gmask], first_cls[gmask]
        next_grasp, predict, _ = self.decode_first_stage(first_grasp, anchors, first_cls)
        m, A = first_cls.shape
        rows = torch.arange(m, device=first_grasp.device)
        ground_gt, ground_score_gt = flat[gmask, :7], flat[gmask, 7:]
        # closest anchor orientation per centre (:99-105): argmin over anchors of 1 - cos(anchor axis, ground axis)
        sim = compute_cos_sim(anchors[:, :, 3:6].reshape(-1, 3),
                              ground_gt[:, None, 3:6].expand(m, A, 3).reshape(-1, 3)).view(m, A)
        ground_8 = torch.sort(sim, dim=1, descending=False)[1][:, 0]
        # class-balanced subset for the anchor classification loss (:111-133)
        counts = torch.bincount(ground_8, minlength=A).cpu().numpy()
        per_anchor = int(counts.min()) if counts.min() > 0 else 1
        picked = []
        for a in range(A):
            members = torch.nonzero(ground_8 == a).view(-1)
            if len(members) == 0:
                continue
            picked.append(members[torch.as_tensor(np.random.choice(len(members), per_anchor, replace=False),
                                                  device=members.device, dtype=torch.long)])
        balanced = torch.cat(picked)
        loss_class = self.criterion_cls(first_cls[balanced], ground_8[balanced].long())
        correct_tuple = ((ground_8 == predict).sum().float(), (ground_8 != predict).sum().float())
        # regression of the ground-truth anchor (:143-162)
        g, tt_gt = first_grasp[rows, ground_8], anchors[rows, ground_8]
        axis = g[:, 3:6] + tt_gt[:, 3:6]
        norm = torch.sqrt((axis * axis).sum(dim=1) + 1e-12).view(-1, 1)
        loss1 = F.smooth_l1_loss(g[:, :3], (ground_gt[:, :3] - tt_gt[:, :3]) / self.radius, reduction='mean')
        loss2 = F.smooth_l1_loss(g[:, 3:6] * norm, ground_gt[:, 3:6] - tt_gt[:, 3:6], reduction='mean')
        loss3 = F.smooth_l1_loss(g[:, 6:7], (ground_gt[:, 6:7] - tt_gt[:, 6:7]) / math.pi, reduction='mean')
        loss4 = F.smooth_l1_loss(g[:, 7:], ground_score_gt, reduction='mean')
        # diagnostics of the predicted anchor against the ground truth (:176-181)
        with torch.no_grad():
            d_center = F.smooth_l1_loss(next_grasp[:, :3], ground_gt[:, :3], reduction='mean')
            d_cos = self._cos_loss(next_grasp[:, 3:6], ground_gt[:, 3:6])
            d_theta = F.smooth_l1_loss(next_grasp[:, 6:7], ground_gt[:, 6:7], reduction='mean')
            d_score = F.smooth_l1_loss(next_grasp[:, 7:], ground_score_gt, reduction='mean')
        next_gt = torch.cat((ground_gt, ground_score_gt), dim=1)
        loss = loss1 * 10 + loss2 * 5 + loss3 + loss4 + loss_class
        loss_tuple = (loss, loss_class.data, loss1.data, loss2.data, loss3.data, loss4.data, d_center, d_cos, d_theta, d_score)
        return next_grasp, loss_tuple, correct_tuple, next_gt, tt_gt, gmask

    # ---- R7: refine decode ----------------------------------------------------------------------------------------
    def compute_loss_refine(self, next_grasp, next_x_cls, next_x_reg, next_gt):
        """(:201-309) final = stage-1 grasp + refine offsets; keep the grasps classified positive, and those that also
        score above the threshold.  With next_gt: a stage-1 grasp counts as positive when it is within 2.5 cm, 60 degrees
        of axis and 60 degrees of angle of its ground truth (:234-243); class-balanced cross-entropy + regression of the
        positives towards their ground truth, and the reference's diagnostics (18-tuple) and confusion counts."""
        final = next_grasp.clone()
        final[:, :3] = final[:, :3] + next_x_reg[:, :3] * self.radius
        final[:, 3:] = final[:, 3:] + next_x_reg[:, 3:]
        predicted = torch.max(next_x_cls, dim=-1)[1]
        positive = predicted == 1
        class_select = torch.nonzero(positive).view(-1)
        score_select = torch.nonzero(positive & (final[:, 7] > self.grasp_score_thre)).view(-1)
        sel_class, sel_score, sel_stage2 = final[class_select].data, final[score_select].data, next_grasp[class_select].data
        if next_gt is None:
            return sel_class, sel_score, sel_stage2, class_select, score_select, (None, None), (None, None, None, None)
        close = torch.sqrt(((next_grasp[:, :3] - next_gt[:, :3]) ** 2).sum(dim=1)) < 0.025
        aligned = compute_cos_sim(next_grasp[:, 3:6], next_gt[:, 3:6]).view(-1) < 0.5
        turned = torch.abs(next_grasp[:, 6] - next_gt[:, 6]) < 1.047
        gt_class = (close & aligned & turned).float()
        gt_1, gt_0 = torch.nonzero(gt_class == 1).view(-1), torch.nonzero(gt_class == 0).view(-1)
        num = min(len(gt_0), len(gt_1))
        zero = next_x_cls.new_zeros(())
        loss = loss_class = l_center = l_r = l_theta = l_score = zero
        diag = [zero] * 12
        if num > 0:
            pick = lambda members: members[torch.as_tensor(np.random.choice(len(members), num, replace=False),
                                                           device=members.device, dtype=torch.long)]
            index = torch.cat((pick(gt_0), pick(gt_1)), dim=-1)
            loss_class = self.criterion_cls(next_x_cls.view(-1, 2)[index], gt_class[index].long())
            l_center = F.smooth_l1_loss(next_x_reg[gt_1, :3], (next_gt[gt_1, :3] - next_grasp[gt_1, :3]) / self.radius)
            l_r = F.smooth_l1_loss(next_x_reg[gt_1, 3:6], next_gt[gt_1, 3:6] - next_grasp[gt_1, 3:6])
            l_theta = F.smooth_l1_loss(next_x_reg[gt_1, 6], next_gt[gt_1, 6] - next_grasp[gt_1, 6])
            l_score = F.smooth_l1_loss(next_x_reg[gt_1, 7:], next_gt[gt_1, 7:] - next_grasp[gt_1, 7:])
            loss = loss_class + l_center + l_r + l_theta + l_score
        if len(class_select) > 0:
            with torch.no_grad():
                def four(pred, sel):
                    return [F.smooth_l1_loss(pred[:, :3], next_gt[sel, :3]), self._cos_loss(pred[:, 3:6], next_gt[sel, 3:6]),
                            F.smooth_l1_loss(pred[:, 6], next_gt[sel, 6]), F.smooth_l1_loss(pred[:, 7:], next_gt[sel, 7:])]
                diag = four(sel_stage2, class_select) + four(sel_class, class_select) + four(sel_score, score_select)
        p1, g1 = predicted.view(-1) == 1, gt_class.view(-1) == 1
        correct = ((g1 & p1).sum().float(), (~g1 & ~p1).sum().float(), (~g1 & p1).sum().float(), (g1 & ~p1).sum().float())
        loss_tuple = (loss, loss_class.data, l_center.data, l_r.data, l_theta.data, l_score) + tuple(diag)
        return sel_class, sel_score, sel_stage2, class_select, score_select, loss_tuple, correct

    def refine_forward(self, pc_group_more_xyz, pc_group_more_index, true_mask, all_feature, group_feature_mp, next_grasp,
                       gripper_params, next_gt=None):
        """(:311-359) closing-box crop of every stage-1 grasp, 64-point feature max-pool, refine head, selection."""
        B, N, C = all_feature.shape
        N_C, N_GM = pc_group_more_index.shape[1], pc_group_more_index.shape[2]
        _, _, inall, gripper_mask = get_gripper_region_transform(
            pc_group_more_xyz[true_mask], pc_group_more_index.view(-1, N_GM)[true_mask], next_grasp, self.gripper_number,
            gripper_params, seed=self.sample_seed, sampler=self._sampler)
        out = (None, None, None, None, None, (None, None), (None, None), next_gt)
        if len(gripper_mask) < 2:
            return out
        cloud = (true_mask // N_C)[gripper_mask]
        flat_index = (inall[gripper_mask].long() + cloud.view(-1, 1) * N).view(1, -1, self.gripper_number)
        pooled = _gather_max(all_feature.contiguous().view(1, B * N, C), flat_index)[0]      # (M', 256)
        centre_half = group_feature_mp.view(-1, 128)[gripper_mask].contiguous()     # the reference's (2M,128) re-view
        next_x_cls, next_x_reg = self.extrat_feature_refine.forward_pooled(pooled, centre_half)
        if next_gt is not None:
            next_gt = next_gt[gripper_mask]
        (select_grasp_class, select_grasp_score, select_grasp_class_stage2, class_select, score_select, loss_refine_tuple,
         correct_refine_tuple) = self.compute_loss_refine(next_grasp[gripper_mask], next_x_cls, next_x_reg, next_gt)
        if next_gt is not None:
            next_gt = next_gt[class_select]
        kept = true_mask[gripper_mask]
        return (select_grasp_class, select_grasp_score, select_grasp_class_stage2, kept[class_select], kept[score_select],
                loss_refine_tuple, correct_refine_tuple, next_gt)

    # ---- module boundary ---------------------------------------------------------------------------------------------
    def forward(self, pc_group, pc_group_more, pc_group_index, pc_group_more_index, center_pc, center_pc_index, pc,
                all_feature, gripper_params, ground_grasp=None, data_path=None):
        """Same arguments and 16-tuple as the reference (:361-434)."""
        B, N_C, N_G, _ = pc_group.shape
        anchors = self._enumerate_anchors(center_pc[:, :, :3].reshape(-1, 3).float())
        pooled = _gather_max(all_feature, pc_group_index.long())                           # (B, N_C, 256): R3
        x_cls, x_reg, mp_center_feature = self.extrat_feature_region.forward_pooled(pooled.view(B * N_C, -1))
        next_grasp, loss_tuple, correct_tuple, next_gt, _, true_mask = self.compute_loss(x_reg, anchors, x_cls, ground_grasp)

        def per_cloud(mask):
            counts = torch.bincount(mask // N_C, minlength=B)
            return [counts[i] for i in range(B)]

        keep_grasp_num_stage2 = per_cloud(true_mask)
        (select_grasp_class, select_grasp_score, select_grasp_class_stage2, final_mask, final_mask_sthre,
         keep_grasp_num_stage3, keep_grasp_num_stage3_score, loss_refine_tuple, correct_refine_tuple, gt) = (None,) * 10
        if self.is_training_refine:
            pc_group_more_xyz = pc_group_more[:, :, :, :6].reshape(B * N_C, -1, 6)
            (select_grasp_class, select_grasp_score, select_grasp_class_stage2, final_mask, final_mask_sthre,
             loss_refine_tuple, correct_refine_tuple, gt) = self.refine_forward(
                pc_group_more_xyz, pc_group_more_index, true_mask, all_feature, mp_center_feature, next_grasp.detach(),
                gripper_params, next_gt)
            if final_mask is not None:
                keep_grasp_num_stage3 = per_cloud(final_mask)
                keep_grasp_num_stage3_score = per_cloud(final_mask_sthre)
            else:
                keep_grasp_num_stage3 = [0 for _ in range(B)]
                keep_grasp_num_stage3_score = [0 for _ in range(B)]
        return (next_grasp.detach(), keep_grasp_num_stage2, true_mask, loss_tuple, correct_tuple, next_gt,
                select_grasp_class, select_grasp_score, select_grasp_class_stage2, keep_grasp_num_stage3,
                keep_grasp_num_stage3_score, final_mask, final_mask_sthre, loss_refine_tuple, correct_refine_tuple, gt)
