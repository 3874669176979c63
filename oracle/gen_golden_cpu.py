"""Generate tests/golden/ref_py_*.npz by importing the REAL reference Python modules from /root/reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference); the fixtures it writes
are committed and are what the tests read.  Run:  python oracle/gen_golden_cpu.py

What is pinned here
  * the module wiring of the reference (concat orders, weight formula, max-pool axis, score head, output
    transposes) -- by running multi_model.score_network.ScoreNetwork / pn2_utils.modules.PointNetSAModule /
    PointnetFPModule unmodified, on CPU, with `pn2_ext` provided by the C oracle (oracle/pn2_oracle.c);
  * that oracle/ref_modules.py (the travelling restatement) reproduces those outputs exactly (asserted below).
The kernels' own arithmetic (FPS / ball query / 3-NN index semantics) is pinned separately against the real
reference CUDA kernels by oracle/gen_golden_gpu.py.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pn2_oracle, ref_modules  # noqa: E402
from regnet_for_3d_grasping_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.modules["pn2_ext"] = pn2_oracle.as_pn2_ext()
    sys.modules["dgcnn_ext"] = pn2_oracle.as_dgcnn_ext()
    sys.path.insert(0, "/root/reference")
    from multi_model.score_network import ScoreNetwork
    from multi_model.utils.pn2_utils import modules as ref_mods
    return ScoreNetwork, ref_mods


def gen_scorenet(ScoreNetwork):
    torch.manual_seed(0)
    B, N = 1, 6144
    pc = torch.from_numpy(synth.batch("table", [11], N))
    sd = ref_modules.random_scorenet_state(seed=3)
    net = ScoreNetwork(training=False).eval()
    missing = net.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert set(net.state_dict().keys()) == set(ref_modules.scorenet_state_shapes().keys())
    with torch.no_grad():
        feat, score, loss = net(pc)
        feat2, score2, dbg = ref_modules.scorenet_forward(sd, pc, pn2_oracle.as_pn2_ext(), keep=True)
    assert loss is None
    assert feat.shape == (B, N, 256) and score.shape == (B, N)
    assert torch.equal(feat, feat2) and torch.equal(score, score2), "ref_modules.py deviates from the reference modules"
    rows = np.arange(0, N, 97)
    np.savez_compressed(
        os.path.join(OUT, "ref_py_scorenet_n6144.npz"),
        meta=np.array("synth.table seed 11, N=6144, B=1; weights ref_modules.random_scorenet_state(seed=3); eval mode"),
        rows=rows.astype(np.int32),
        all_feature_rows=feat[0, rows].numpy(),
        score=score[0].numpy(),
        fps0=dbg["fps0"][0].numpy().astype(np.int32), fps1=dbg["fps1"][0].numpy().astype(np.int32),
        fps2=dbg["fps2"][0].numpy().astype(np.int32),
        bqcnt0=dbg["bqcnt0"][0].numpy().astype(np.int16), bqcnt1=dbg["bqcnt1"][0].numpy().astype(np.int16),
        bqcnt2=dbg["bqcnt2"][0].numpy().astype(np.int16),
        bq0_sum=dbg["bq0"][0].sum(1).numpy(), bq1_sum=dbg["bq1"][0].sum(1).numpy(), bq2_sum=dbg["bq2"][0].sum(1).numpy(),
        nn2=dbg["nn2"][0, ::13].numpy().astype(np.int32),
        sa2_rows=dbg["sa2"][0, :, ::16].numpy(),
    )
    # train-mode loss path: MSE against pc_score (score_network.py:19-29,50-51); dropout makes the score
    # random in train mode, so only eval-with-loss is pinned: is_training=True but module in eval()
    net.is_training = True
    tgt = torch.from_numpy(synth.scores_like_dataset(5, B, N))
    with torch.no_grad():
        _, s3, loss3 = net(pc, tgt)
    np.savez_compressed(os.path.join(OUT, "ref_py_scorenet_loss.npz"), loss=loss3.numpy(),
                        meta=np.array("same input; pc_score = synth.scores_like_dataset(5,1,6144); MSELoss mean"))
    print("scorenet golden ok: loss", float(loss3))


def gen_modules(ref_mods):
    """Small SA / FP module fixtures with the full tensors stored (weights included)."""
    g = torch.Generator().manual_seed(7)
    B, N, M, K, C = 2, 512, 128, 16, 5
    pc = torch.from_numpy(synth.batch("cube", [21, 22], N))
    xyz = pc[:, :, :3].permute(0, 2, 1)
    feat = torch.randn(B, C, N, generator=g)
    sa = ref_mods.PointNetSAModule(in_channels=C, mlp_channels=(16, 24), num_centroids=M, radius=0.15,
                                   num_neighbours=K, use_xyz=True).eval()
    fp = ref_mods.PointnetFPModule(in_channels=24 + C, mlp_channels=(20, 12), num_neighbors=3).eval()
    with torch.no_grad():
        for mod in list(sa.modules()) + list(fp.modules()):
            if isinstance(mod, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) + 0.5)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) + 0.5)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
        new_xyz, new_feat = sa(xyz, feat)
        up = fp(xyz, new_xyz, feat, new_feat)
        gi = ref_mods._F.farthest_point_sample(xyz, M)
        bi, bc = ref_mods._F.ball_query(xyz, new_xyz, 0.15, K)
        ni, nd = ref_mods._F.search_nn_distance(xyz, new_xyz, 3)
    out = {"pc": pc.numpy(), "feat": feat.numpy(), "new_xyz": new_xyz.numpy(), "new_feat": new_feat.numpy(),
           "up": up.numpy(), "fps": gi.numpy().astype(np.int32), "bq": bi.numpy().astype(np.int32),
           "bqcnt": bc.numpy().astype(np.int32), "nn": ni.numpy().astype(np.int32), "nnd": nd.numpy(),
           "meta": np.array("PointNetSAModule(C=5,(16,24),M=128,r=.15,K=16) + PointnetFPModule(29,(20,12),3), eval")}
    for k, v in sa.state_dict().items():
        out["sa." + k] = v.numpy()
    for k, v in fp.state_dict().items():
        out["fp." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "ref_py_modules_small.npz"), **out)
    print("module golden ok", new_feat.shape, up.shape)

    # autograd through the reference's Function wrappers (function.py:84-107, 146-172): gradient fixtures
    x = torch.randn(B, C, N, generator=g, requires_grad=True)
    grouped = ref_mods._F.group_points(x, bi)
    wsum = torch.randn(grouped.shape, generator=g)
    (grouped * wsum).sum().backward()
    g_group = x.grad.clone()
    s = torch.randn(B, 7, M, generator=g, requires_grad=True)
    inv = 1.0 / torch.clamp(nd, min=1e-10)
    w = inv / inv.sum(2, keepdim=True)
    it = ref_mods._F.feature_interpolate(s, ni, w)
    wsum2 = torch.randn(it.shape, generator=g)
    (it * wsum2).sum().backward()
    np.savez_compressed(os.path.join(OUT, "ref_py_grads_small.npz"), x=x.detach().numpy(), wsum=wsum.numpy(),
                        g_group=g_group.numpy(), s=s.detach().numpy(), w=w.numpy(), wsum2=wsum2.numpy(),
                        interp=it.detach().numpy(), g_interp=s.grad.numpy(),
                        meta=np.array("uses bq / nn / nnd of ref_py_modules_small.npz"))
    print("grad golden ok")


def gen_heads():
    """Region / refine heads (utils/pointnet2.py:123-254) of the reference on seeded inputs, eval mode."""
    from multi_model.utils.pointnet2 import PointNet2Refine, PointNet2TwoStage
    g = torch.Generator().manual_seed(17)

    from regnet_for_3d_grasping_b200.weights import seeded_state_like

    def randomize(mod):
        mod.load_state_dict(seeded_state_like(mod.state_dict(), seed=17), strict=True)

    two = PointNet2TwoStage(num_points=16, input_chann=6, k_cls=4, k_reg=40, k_reg_theta=4).eval()
    ref = PointNet2Refine(num_points=8, input_chann=6, k_cls=2, k_reg=10).eval()
    randomize(two)
    randomize(ref)
    x = torch.randn(12, 256, 16, generator=g)
    gf = torch.randn(7, 256, 8, generator=g)
    grp = torch.randn(7, 128, generator=g)
    with torch.no_grad():
        cls, reg, mp = two(x, None)
        rcls, rreg = ref(gf, grp)
    out = {"x": x.numpy(), "gf": gf.numpy(), "grp": grp.numpy(), "cls": cls.numpy(), "reg": reg.numpy(),
           "mp": mp.numpy(), "rcls": rcls.numpy(), "rreg": rreg.numpy(),
           "meta": np.array("PointNet2TwoStage(16,6,4,40,4) / PointNet2Refine(8,6,2,10), eval; weights = weights.seeded_state_like(state_dict, seed=17)")}
    np.savez_compressed(os.path.join(OUT, "ref_py_heads.npz"), **out)
    print("heads golden ok", cls.shape, reg.shape, rcls.shape)


def deterministic_choice(n, k, replace=True):
    """Stand-in for np.random.choice inside the reference while the region-net fixture is generated: the reference's
    picks come from a wall-clock-seeded global RNG and cannot be reproduced, a fixed rule can.  Without replacement:
    the first k; with replacement: (7 i + 3) mod n.  tests/helpers.py applies the same rule to the product's mask."""
    return np.arange(k) if not replace else (np.arange(k) * 7 + 3) % n


def region_net_inputs(seed=5, B=2, N=900, N_C=6, N_G=16, N_GM=96):
    """Seeded inputs of GripperRegionNetwork.forward with real geometry: centres on the cloud, groups = nearest points."""
    g = torch.Generator().manual_seed(seed)
    pc = torch.from_numpy(synth.batch("table", [31, 32][:B], N))
    # shrink the scene around its centroid: ~15 points per cm^2, so that the 3 x 8 x 1 cm closing boxes of randomly
    # oriented grasps hold anything from 0 to a few dozen points (accept / reject / both sampling branches all occur)
    mean = pc[:, :, :3].mean(dim=1, keepdim=True)
    pc[:, :, :3] = (pc[:, :, :3] - mean) * 0.25 + mean
    all_feature = torch.randn(B, N, 256, generator=g)
    center_idx = torch.stack([torch.randperm(N, generator=g)[:N_C] for _ in range(B)])
    center_pc = torch.stack([pc[b, center_idx[b]] for b in range(B)])
    d = torch.cdist(center_pc[:, :, :3], pc[:, :, :3])                      # (B, N_C, N)
    more_idx = d.topk(N_GM, dim=2, largest=False)[1]
    grp_idx = more_idx[:, :, :N_G].contiguous()
    gather = lambda idx: torch.stack([pc[b][idx[b]] for b in range(B)])
    return dict(pc=pc, all_feature=all_feature, center_pc=center_pc, center_pc_index=center_idx, pc_group_index=grp_idx,
                pc_group=gather(grp_idx), pc_group_more_index=more_idx, pc_group_more=gather(more_idx))


def gen_region_net(weight_seed=46, save=True):
    """multi_model/gripper_region_network.py, inference call (ground_grasp=None), run UNMODIFIED on CPU: its
    unconditional `.cuda()` calls are made no-ops and np.random.choice is replaced by deterministic_choice."""
    import multi_model.gripper_region_network as ref
    from regnet_for_3d_grasping_b200.weights import seeded_state_like
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref.np.random.choice = deterministic_choice
    net = ref.GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                                   reg_channel=10).eval()
    net.load_state_dict(seeded_state_like(net.state_dict(), seed=weight_seed), strict=True)
    inp = region_net_inputs()
    params = [0.08, 0.010, 0.06]
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = net(inp["pc_group"], inp["pc_group_more"], inp["pc_group_index"], inp["pc_group_more_index"], inp["center_pc"],
                  inp["center_pc_index"], inp["pc"], inp["all_feature"], params)
        M, NGM = 12, inp["pc_group_more"].shape[2]
        gp, gi, ginall, gmask = ref.get_gripper_region_transform(inp["pc_group_more"].view(M, NGM, 6),
                                                                 inp["pc_group_more_index"].view(M, NGM), out[0], 8, params)
    (next_grasp, keep2, true_mask, loss_tuple, correct_tuple, next_gt, sel_class, sel_score, sel_stage2, keep3, keep3s,
     final_mask, final_mask_sthre, loss_refine, correct_refine, gt) = out
    assert loss_tuple == (None, None) and next_gt is None and gt is None
    pcs_t_mask = [int((gi[m] >= 0).any()) for m in range(M)]
    print("accepted:", pcs_t_mask, "gripper_mask", gmask.tolist())
    if not save:
        return len(gmask), (0 if final_mask is None else len(final_mask)), (0 if final_mask_sthre is None else len(final_mask_sthre))
    assert final_mask is not None and len(final_mask) > 0, "fixture should exercise the refine stage"
    save = {k: v.numpy() for k, v in inp.items()}
    save.update(next_grasp=next_grasp.numpy(), keep2=np.array([int(k) for k in keep2]), true_mask=true_mask.numpy(),
                sel_class=sel_class.numpy(), sel_score=sel_score.numpy(), sel_stage2=sel_stage2.numpy(),
                keep3=np.array([int(k) for k in keep3]), keep3s=np.array([int(k) for k in keep3s]),
                final_mask=final_mask.numpy(), final_mask_sthre=final_mask_sthre.numpy(),
                gripper_pc=gp.numpy(), gripper_pc_index=gi.numpy(), gripper_pc_index_inall=ginall.numpy(),
                gripper_mask=gmask.numpy(), params=np.array(params),
                meta=np.array("GripperRegionNetwork(True,16,8,0.4,0.06,10).eval(), weights seeded_state_like(seed=weight_seed); "
                              "inputs gen_golden_cpu.region_net_inputs(); np.random.choice -> deterministic_choice"),
                weight_seed=np.array(weight_seed))
    np.savez_compressed(os.path.join(OUT, "ref_py_region_net.npz"), **save)
    print("region-net golden ok: accepted grasps", len(gmask), "of", M, "; positives", len(final_mask), "; above threshold",
          len(final_mask_sthre))


def region_loss_inputs(seed=3, B=2, NC=40, A=4):
    """Seeded inputs of GripperRegionNetwork.compute_loss / compute_loss_refine (training call): regressions, anchor
    scores, and a ground truth (B, NC, 10) = (centre, axis, theta, 3 scores) with two centres that have no grasp (-1)."""
    g = torch.Generator().manual_seed(seed)
    M = B * NC
    first_grasp = torch.randn(M, A, 10, generator=g) * 0.3
    first_grasp[:, :, 7:] = torch.rand(M, A, 3, generator=g)
    centers = torch.rand(M, 3, generator=g)
    first_cls = torch.randn(M, A, generator=g)
    ground = torch.cat([centers.view(B, NC, 3) + torch.randn(B, NC, 3, generator=g) * 0.01,
                        torch.nn.functional.normalize(torch.randn(B, NC, 3, generator=g), dim=-1),
                        (torch.rand(B, NC, 1, generator=g) - 0.5) * 3, torch.rand(B, NC, 3, generator=g)], dim=-1)
    ground[0, 3, -1] = -1
    ground[1, 7, -1] = -1
    m = 50
    next_grasp = torch.cat([torch.rand(m, 3, generator=g), torch.nn.functional.normalize(torch.randn(m, 3, generator=g), dim=-1),
                            (torch.rand(m, 1, generator=g) - 0.5) * 3, torch.rand(m, 3, generator=g)], dim=-1)
    next_gt = next_grasp.clone()
    next_gt[:, :3] += torch.randn(m, 3, generator=g) * 0.02          # some within 2.5 cm, some not
    next_gt[:, 3:6] = torch.nn.functional.normalize(next_gt[:, 3:6] + torch.randn(m, 3, generator=g) * 0.5, dim=-1)
    next_gt[:, 6] += torch.randn(m, generator=g) * 0.8
    next_x_cls = torch.randn(m, 2, generator=g)
    next_x_reg = torch.randn(m, 10, generator=g) * 0.1
    return dict(first_grasp=first_grasp, centers=centers, first_cls=first_cls, ground=ground, next_grasp=next_grasp,
                next_gt=next_gt, next_x_cls=next_x_cls, next_x_reg=next_x_reg)


def gen_region_losses():
    """The TRAINING branches of multi_model/gripper_region_network.py (compute_loss with a ground truth, :92-199, and
    compute_loss_refine with next_gt, :217-309) run on CPU from the reference's own file.  Two shims, both stated here:
    `.cuda()` is a no-op and np.random.choice is deterministic_choice (as for the inference fixture); and
    nn.CosineEmbeddingLoss receives a 1-D all-ones target of the inputs' length -- torch >= 1.10 rejects the (n, 1)
    target the reference passes (once even with a mismatched n, :283), torch 1.8 (the reference's pin) broadcast it,
    which for an all-ones target gives mean(1 - cos) whatever its length."""
    import contextlib
    import io
    import multi_model.gripper_region_network as ref
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref.np.random.choice = deterministic_choice
    net = ref.GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                                   reg_channel=10).eval()
    class _Cos(torch.nn.CosineEmbeddingLoss):
        def forward(self, a, b, y):
            return super().forward(a, b, a.new_ones(a.shape[0]))   # what the (n, 1) ones target meant under torch 1.8
    net.criterion_cos = _Cos(reduction="mean")
    inp = region_loss_inputs()
    anchors = net._enumerate_anchors(inp["centers"])
    with contextlib.redirect_stdout(io.StringIO()):
        next_grasp, loss_tuple, correct_tuple, next_gt, tt_gt, gmask = net.compute_loss(
            inp["first_grasp"].clone(), anchors, inp["first_cls"].clone(), inp["ground"].clone())
        (sel_class, sel_score, sel_stage2, class_select, score_select, loss_refine, correct_refine) = net.compute_loss_refine(
            inp["next_grasp"].clone(), inp["next_x_cls"].clone(), inp["next_x_reg"].clone(), inp["next_gt"].clone())
    out = {k: v.numpy() for k, v in inp.items()}
    out.update(l_next_grasp=next_grasp.numpy(), l_loss=np.array([float(x) for x in loss_tuple]),
               l_correct=np.array([float(x) for x in correct_tuple]), l_next_gt=next_gt.numpy(), l_tt_gt=tt_gt.numpy(),
               l_gmask=gmask.numpy(), r_sel_class=sel_class.numpy(), r_sel_score=sel_score.numpy(),
               r_sel_stage2=sel_stage2.numpy(), r_class_select=class_select.numpy(), r_score_select=score_select.numpy(),
               r_loss=np.array([float(x) for x in loss_refine]), r_correct=np.array([float(x) for x in correct_refine]),
               meta=np.array("GripperRegionNetwork(True,16,8,0.4,0.06,10): compute_loss / compute_loss_refine with ground "
                             "truth; inputs gen_golden_cpu.region_loss_inputs(); np.random.choice -> deterministic_choice"))
    np.savez_compressed(os.path.join(OUT, "ref_py_region_losses.npz"), **out)
    print("region-loss golden ok: stage-1 loss", out["l_loss"][0], "acc", out["l_correct"], "; refine loss", out["r_loss"][0],
          "TP/TN/FP/FN", out["r_correct"])


def region_net_ground(inp, seed=23):
    """A ground truth for region_net_inputs(): (B, N_C, 10) = grasp near each centre (centre, axis, theta, 3 scores); one
    centre without a grasp."""
    g = torch.Generator().manual_seed(seed)
    c = inp["center_pc"][:, :, :3]
    B, NC = c.shape[:2]
    ground = torch.cat([c + torch.randn(B, NC, 3, generator=g) * 0.004,
                        torch.nn.functional.normalize(torch.randn(B, NC, 3, generator=g), dim=-1),
                        (torch.rand(B, NC, 1, generator=g) - 0.5) * 3, torch.rand(B, NC, 3, generator=g)], dim=-1)
    ground[1, 2, -1] = -1
    return ground


def gen_region_net_train(weight_seed=46):
    """The TRAINING call of the reference's GripperRegionNetwork.forward (ground_grasp given), eval-mode BatchNorm so that
    the fixture does not depend on batch statistics; shims as in gen_region_losses()."""
    import contextlib
    import io
    import multi_model.gripper_region_network as ref
    from regnet_for_3d_grasping_b200.weights import seeded_state_like
    torch.Tensor.cuda = lambda self, *a, **k: self
    ref.np.random.choice = deterministic_choice
    net = ref.GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                                   reg_channel=10).eval()
    net.load_state_dict(seeded_state_like(net.state_dict(), seed=weight_seed), strict=True)

    class _Cos(torch.nn.CosineEmbeddingLoss):
        def forward(self, a, b, y):
            return super().forward(a, b, a.new_ones(a.shape[0]))
    net.criterion_cos = _Cos(reduction="mean")
    inp = region_net_inputs()
    ground = region_net_ground(inp)
    params = [0.08, 0.010, 0.06]
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = net(inp["pc_group"], inp["pc_group_more"], inp["pc_group_index"], inp["pc_group_more_index"], inp["center_pc"],
                  inp["center_pc_index"], inp["pc"], inp["all_feature"], params, ground)
    (next_grasp, keep2, true_mask, loss_tuple, correct_tuple, next_gt, sel_class, sel_score, sel_stage2, keep3, keep3s,
     final_mask, final_mask_sthre, loss_refine, correct_refine, gt) = out
    assert final_mask is not None and loss_refine[0] is not None, "fixture should exercise the refine losses"
    f = lambda tup: np.array([float(x) for x in tup])
    np.savez_compressed(os.path.join(OUT, "ref_py_region_net_train.npz"), ground=ground.numpy(), next_grasp=next_grasp.numpy(),
                        keep2=np.array([int(k) for k in keep2]), true_mask=true_mask.numpy(), loss=f(loss_tuple),
                        correct=f(correct_tuple), next_gt=next_gt.numpy(), sel_class=sel_class.numpy(),
                        sel_score=sel_score.numpy(), sel_stage2=sel_stage2.numpy(), keep3=np.array([int(k) for k in keep3]),
                        keep3s=np.array([int(k) for k in keep3s]), final_mask=final_mask.numpy(),
                        final_mask_sthre=final_mask_sthre.numpy(), loss_refine=f(loss_refine), correct_refine=f(correct_refine),
                        gt=gt.numpy(), weight_seed=np.array(weight_seed),
                        meta=np.array("training call of GripperRegionNetwork(True,16,8,0.4,0.06,10).eval() on "
                                      "region_net_inputs() + region_net_ground(); np.random.choice -> deterministic_choice"))
    print("region-net training golden ok: loss", f(loss_tuple)[0], "refine loss", f(loss_refine)[0], "kept", len(true_mask),
          "final", len(final_mask), "gt", tuple(gt.shape))


def gen_center_grasp():
    """dataset_utils/get_regiondataset.py:_get_center_grasp (+ _transform_grasp), the reference's own functions on CPU,
    over two synthetic scene files (synth.scene_grasps).  `open3d`, which that module imports and these functions never
    touch, is stubbed with an empty module; `.cuda()` is a no-op."""
    import contextlib
    import io
    import tempfile
    import types
    sys.modules.setdefault("open3d", types.ModuleType("open3d"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    import dataset_utils.get_regiondataset as ref
    inp = region_net_inputs(N_C=24)
    with tempfile.TemporaryDirectory() as tmp:
        paths = [synth.write_scene_file(os.path.join(tmp, f"scene{b}.p"), 70 + b, inp["pc"][b].numpy(), n_grasps=10, hit_frac=0.5) for b in range(2)]
        with contextlib.redirect_stdout(io.StringIO()):
            labels = ref._get_center_grasp(inp["center_pc_index"], inp["center_pc"], paths, 0.06, True)
            frames = ref._get_center_grasp(inp["center_pc_index"], inp["center_pc"], paths, 0.06, False)
    found = int((labels[:, :, 7] != -1).sum())
    assert 5 < found < labels.shape[0] * labels.shape[1], "fixture should have centres with and without a grasp"
    np.savez_compressed(os.path.join(OUT, "ref_py_center_grasp.npz"), center_pc=inp["center_pc"].numpy(),
                        center_pc_index=inp["center_pc_index"].numpy(), pc=inp["pc"].numpy(), labels=labels.numpy(),
                        frames=frames.numpy(),
                        meta=np.array("_get_center_grasp(center_pc_index, center_pc, [scene0, scene1], depth=0.06) with "
                                      "use_theta True / False; scenes = synth.write_scene_file(seed 70 + b, pc[b], n_grasps=10, hit_frac=0.5); "
                                      "inputs region_net_inputs(N_C=24)"))
    print("centre-grasp golden ok:", tuple(labels.shape), "labelled centres", found)


def eval_test_inputs(seed=41, N=3000, M=400):
    """A table-top view cloud and grasps placed on it (some upright over the objects, some tilted into the table, some
    floating): (points (N,3), grasps (M,8), table_height, depth, width)."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.from_numpy(synth.batch("table", [seed], N))[0, :, :3]
    anchor = pts[torch.randint(0, N, (M,), generator=g)]
    centre = anchor + torch.randn(M, 3, generator=g) * 0.01 + torch.tensor([0.0, 0.0, 0.03]) * torch.rand(M, 1, generator=g)
    axis = torch.nn.functional.normalize(torch.randn(M, 3, generator=g) * torch.tensor([1.0, 1.0, 0.2]), dim=-1)
    angle = (torch.rand(M, 1, generator=g) - 0.5) * 3.0
    grasp = torch.cat([centre, axis, angle, torch.rand(M, 1, generator=g)], dim=1)
    return pts, grasp, 0.75, 0.06, 0.08


def gen_eval_test():
    """dataset_utils/eval_score/eval.py:eval_test, the reference's own EvalDataTest.run_collision_view on CPU (gpu=-1),
    on top of the open3d / transforms3d stand-ins of dropin/ (the reference estimates normals in the constructor and never
    uses them in this filter)."""
    import contextlib
    import io
    sys.path.insert(0, os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin"))
    for name in ("open3d", "transforms3d"):
        sys.modules.pop(name, None)
    import importlib
    importlib.import_module("open3d")
    importlib.import_module("transforms3d")
    from dataset_utils.eval_score.eval_utils.evaluation_data_generator import EvalDataTest
    pts, grasp, table_height, depth, width = eval_test_inputs()
    with contextlib.redirect_stdout(io.StringIO()):
        ev = EvalDataTest(pts.numpy(), grasp.clone(), None, table_height, depth, width, -1)
        kept = ev.run_collision_view()
        idx = ev.baseline_frame_index[:ev.valid_grasp].long()
    assert 10 < len(idx) < len(grasp) - 10, f"fixture should keep some grasps and reject some (kept {len(idx)})"
    np.savez_compressed(os.path.join(OUT, "ref_py_eval_test.npz"), kept_index=idx.numpy(), kept=kept.numpy(),
                        frame=ev.frame.numpy(), meta=np.array("EvalDataTest(points, grasp, None, 0.75, 0.06, 0.08, -1)"
                        ".run_collision_view() on gen_golden_cpu.eval_test_inputs()"))
    print("eval_test golden ok: kept", len(idx), "of", len(grasp))


def eval_validate_inputs(seed=47, N=20000, N2=60000, M=4000):
    """View cloud, a denser scene cloud of the same synthetic scene with random unit normals (the score only uses |n_y|
    in the gripper frame), and top-down grasps over the objects (approach ~ -z, random yaw, small tilt): some straddle an
    object, most hit something with a finger."""
    g = torch.Generator().manual_seed(seed)
    pts = torch.from_numpy(synth.batch("table", [seed], N))[0, :, :3]
    scene = torch.from_numpy(synth.batch("table", [seed], N2))[0, :, :3]
    pts[:, 2], scene[:, 2] = 1.5 - pts[:, 2], 1.5 - scene[:, 2]      # synth's objects lie towards the camera (smaller z):
    normals = torch.nn.functional.normalize(torch.randn(N2, 3, generator=g) * torch.tensor([0.6, 0.6, 1.0]), dim=-1)
    obj = pts[pts[:, 2] > 0.775]                                     # mirrored about the table plane they stand ON it
    anchor = obj[torch.randint(0, len(obj), (M,), generator=g)]
    centre = anchor + torch.cat([torch.randn(M, 2, generator=g) * 0.01, torch.rand(M, 1, generator=g) * 0.04], dim=1)
    yaw = torch.rand(M, generator=g) * 6.2832
    axis = torch.stack([torch.cos(yaw), torch.sin(yaw), torch.randn(M, generator=g) * 0.05], dim=1)
    angle = -1.5708 + torch.randn(M, 1, generator=g) * 0.2
    grasp = torch.cat([centre, torch.nn.functional.normalize(axis, dim=-1), angle, torch.rand(M, 1, generator=g)], dim=1)
    data = dict(view_cloud=pts.numpy(), scene_cloud=scene.numpy().astype(np.float64),
                scene_normal=normals.numpy().astype(np.float64))
    return data, grasp, 0.75, 0.06, 0.08


def gen_eval_validate():
    """dataset_utils/eval_score/eval.py:eval_validate, the reference's own EvalDataValidate.run_collision on CPU
    (gpu=-1; torch.cuda.is_available() is False here, so its scene tensors stay on the CPU too), on the stand-ins."""
    import contextlib
    import importlib
    import io
    sys.path.insert(0, os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin"))
    for name in ("open3d", "transforms3d"):
        sys.modules.pop(name, None)
    importlib.import_module("open3d")
    importlib.import_module("transforms3d")
    from dataset_utils.eval_score.eval_utils.evaluation_data_generator import EvalDataValidate
    data, grasp, table_height, depth, width = eval_validate_inputs()
    with contextlib.redirect_stdout(io.StringIO()):
        ev = EvalDataValidate(data, grasp.clone(), 0, table_height, depth, width, -1)
        vgr, score, n_view, g_view, g_scene = ev.run_collision()
    assert 5 < vgr < n_view < len(grasp), f"fixture should exercise both filters (view {n_view}, scene {vgr})"
    np.savez_compressed(os.path.join(OUT, "ref_py_eval_validate.npz"), vgr=np.array(vgr), score=np.array(score),
                        n_view=np.array(n_view), grasp_view=g_view.numpy(), grasp_scene=g_scene.numpy(),
                        antipodal=ev.antipodal_score.numpy(),
                        meta=np.array("EvalDataValidate(data, grasp, 0, 0.75, 0.06, 0.08, -1).run_collision() on "
                                      "gen_golden_cpu.eval_validate_inputs()"))
    print("eval_validate golden ok: view", n_view, "scene", vgr, "score", score)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ScoreNetwork, ref_mods = import_reference()
    if "heads" in sys.argv:
        gen_heads()
        sys.exit(0)
    if "region_net" in sys.argv:
        gen_region_net()
        sys.exit(0)
    if "eval_validate" in sys.argv:
        gen_eval_validate()
        sys.exit(0)
    if "eval_test" in sys.argv:
        gen_eval_test()
        sys.exit(0)
    if "center_grasp" in sys.argv:
        gen_center_grasp()
        sys.exit(0)
    if "region_net_train" in sys.argv:
        gen_region_net_train()
        sys.exit(0)
    if "region_losses" in sys.argv:
        gen_region_losses()
        sys.exit(0)
    gen_modules(ref_mods)
    gen_scorenet(ScoreNetwork)
    gen_heads()
    gen_region_net()
    gen_region_losses()
    gen_region_net_train()
    gen_center_grasp()
    gen_eval_test()
    gen_eval_validate()
