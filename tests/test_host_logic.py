"""CPU suite, part 3: host-side mirror of the reference interface (modules, state dict, BN folding).

The CUDA operators cannot run here, so the module wiring is exercised with the C oracle patched in under
`function.pn2_ext` -- the oracle is the checker's stand-in for the kernels, the modules under test are the
product's.  Outputs are compared with fixtures produced by the reference's own modules."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import golden


@pytest.fixture()
def oracle_ops(oracle, monkeypatch):
    from regnet_for_3d_grasping_b200 import function
    monkeypatch.setattr(function, "pn2_ext", oracle.as_pn2_ext())
    return oracle


def test_state_dict_schema_matches_reference():
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
    net = ScoreNetwork(training=False)
    sd = net.state_dict()
    want = ref_modules.scorenet_state_shapes()
    assert set(sd.keys()) == set(want.keys())
    assert len(sd) == 127
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    assert sum(p.numel() for p in net.parameters()) == 5542531


def test_sa_fp_modules_match_reference_outputs(oracle_ops):
    ref = golden("ref_py_modules_small.npz")
    from regnet_for_3d_grasping_b200.modules import PointNetSAModule, PointnetFPModule
    sa = PointNetSAModule(in_channels=5, mlp_channels=(16, 24), num_centroids=128, radius=0.15, num_neighbours=16,
                          use_xyz=True).eval()
    fp = PointnetFPModule(in_channels=29, mlp_channels=(20, 12), num_neighbors=3).eval()
    sa.load_state_dict({k[3:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith("sa.")})
    fp.load_state_dict({k[3:]: torch.from_numpy(ref[k]) for k in ref.files if k.startswith("fp.")})
    pc = torch.from_numpy(ref["pc"])
    xyz = pc[:, :, :3].permute(0, 2, 1)
    feat = torch.from_numpy(ref["feat"])
    with torch.no_grad():
        new_xyz, new_feat = sa(xyz, feat)
        up = fp(xyz, new_xyz, feat, new_feat)
    assert np.array_equal(new_xyz.numpy(), ref["new_xyz"])
    np.testing.assert_allclose(new_feat.numpy(), ref["new_feat"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(up.numpy(), ref["up"], rtol=1e-6, atol=1e-6)


def test_autograd_wrappers_match_reference_gradients(oracle_ops):
    ref = golden("ref_py_grads_small.npz")
    mods = golden("ref_py_modules_small.npz")
    from regnet_for_3d_grasping_b200 import function as F
    bq = torch.from_numpy(mods["bq"]).long()
    nn = torch.from_numpy(mods["nn"]).long()
    x = torch.from_numpy(ref["x"]).requires_grad_(True)
    (F.group_points(x, bq) * torch.from_numpy(ref["wsum"])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), ref["g_group"], rtol=1e-5, atol=1e-5)
    s = torch.from_numpy(ref["s"]).requires_grad_(True)
    it = F.feature_interpolate(s, nn, torch.from_numpy(ref["w"]))
    np.testing.assert_allclose(it.detach().numpy(), ref["interp"], rtol=1e-6, atol=1e-6)
    (it * torch.from_numpy(ref["wsum2"])).sum().backward()
    np.testing.assert_allclose(s.grad.numpy(), ref["g_interp"], rtol=1e-5, atol=1e-5)
    # index-producing ops are non-differentiable (function.py:46-48,76-78,131-133)
    pts = torch.rand(1, 3, 32, requires_grad=True)
    assert not F.farthest_point_sample(pts, 8).requires_grad
    assert not F.ball_query(pts, pts[:, :, :4], 0.5, 4)[0].requires_grad
    assert not F.search_nn_distance(pts, pts[:, :, :8], 3)[0].requires_grad


def test_scorenetwork_train_path_matches_reference(oracle_ops):
    """Op-by-op path of the drop-in ScoreNetwork (module in eval() so BN/dropout are deterministic) against the
    fixture from the reference's ScoreNetwork, including the loss."""
    ref = golden("ref_py_scorenet_n6144.npz")
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
    net = ScoreNetwork(training=True).eval()
    net.load_state_dict(ref_modules.random_scorenet_state(seed=3))
    pc = torch.from_numpy(synth.batch("table", [11], 6144))
    tgt = torch.from_numpy(synth.scores_like_dataset(5, 1, 6144))
    with torch.no_grad():
        feat, score, loss = net(pc, tgt)      # CPU tensors -> not fusable -> module path
    assert feat.shape == (1, 6144, 256) and score.shape == (1, 6144)
    np.testing.assert_allclose(feat[0, ref["rows"]].numpy(), ref["all_feature_rows"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(score[0].numpy(), ref["score"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(loss.numpy(), golden("ref_py_scorenet_loss.npz")["loss"], rtol=1e-5)
    net.is_training = False
    with torch.no_grad():
        assert net(pc, tgt)[2] is None


def test_fold_bn_equals_batchnorm_eval():
    from regnet_for_3d_grasping_b200.nn_layers import Conv1d
    from regnet_for_3d_grasping_b200.scorenet import fold_bn
    torch.manual_seed(0)
    blk = Conv1d(7, 5, 1).eval()
    with torch.no_grad():
        blk.bn.running_mean.normal_(0, 0.3)
        blk.bn.running_var.uniform_(0.5, 2.0)
        blk.bn.weight.uniform_(0.5, 1.5)
        blk.bn.bias.normal_(0, 0.2)
        x = torch.randn(3, 7, 11)
        want = blk(x)
        w, scale, shift = fold_bn({"p." + k: v for k, v in blk.state_dict().items()}, "p")
        got = torch.relu(torch.einsum("oc,bcn->bon", w, x) * scale[None, :, None] + shift[None, :, None])
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_synthetic_clouds_are_seeded_and_shaped():
    from regnet_for_3d_grasping_b200 import synth
    a, b = synth.table_scene(5, 25600), synth.table_scene(5, 25600)
    assert a.shape == (25600, 6) and a.dtype == np.float32 and np.array_equal(a, b)
    assert not np.array_equal(a, synth.table_scene(6, 25600))
    assert 0.5 < a[:, 2].mean() < 0.8 and a[:, 3:].min() >= 0 and a[:, 3:].max() <= 1
    lat = synth.lattice(1, 512)
    assert len(np.unique(lat[:, :3], axis=0)) < 512   # duplicates exist: tie paths are exercised


def test_region_heads_match_reference_outputs():
    """PointNet2TwoStage / PointNet2Refine mirrors (rows R4, R7): same state-dict keys, same outputs as the fixture
    produced by the reference's classes; forward_pooled (the gather-max entry point) equals forward."""
    ref = golden("ref_py_heads.npz")
    from regnet_for_3d_grasping_b200.region_heads import PointNet2Refine, PointNet2TwoStage
    two = PointNet2TwoStage(num_points=16, input_chann=6, k_cls=4, k_reg=40, k_reg_theta=4).eval()
    rf = PointNet2Refine(num_points=8, input_chann=6, k_cls=2, k_reg=10).eval()
    from regnet_for_3d_grasping_b200.weights import seeded_state_like
    two.load_state_dict(seeded_state_like(two.state_dict(), seed=17), strict=True)   # same keys => same weights as the
    rf.load_state_dict(seeded_state_like(rf.state_dict(), seed=17), strict=True)      # fixture generator used
    x, gf, grp = (torch.from_numpy(ref[k]) for k in ("x", "gf", "grp"))
    with torch.no_grad():
        cls, reg, mp = two(x, None)
        cls2, reg2, mp2 = two.forward_pooled(x.max(dim=2)[0])
        rcls, rreg = rf(gf, grp)
        rcls2, rreg2 = rf.forward_pooled(gf.max(dim=2)[0], grp)
    np.testing.assert_allclose(cls.numpy(), ref["cls"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(reg.numpy(), ref["reg"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mp.numpy(), ref["mp"], rtol=0, atol=0)
    np.testing.assert_allclose(rcls.numpy(), ref["rcls"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rreg.numpy(), ref["rreg"], rtol=1e-5, atol=1e-6)
    assert torch.equal(cls, cls2) and torch.equal(reg, reg2) and torch.equal(mp, mp2)
    assert torch.equal(rcls, rcls2) and torch.equal(rreg, rreg2)
    assert tuple(reg.shape) == (12, 4, 10) and (reg[:, :, 7:] > 0).all() and (reg[:, :, 7:] < 1).all()


def test_region_network_state_dict_schema():
    """GripperRegionNetwork mirror: Appendix B keys and parameter count (1 524 396), templates rounded through fp16, and
    the drop-in import path resolves to it."""
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    from regnet_for_3d_grasping_b200.gripper_region_network import GripperRegionNetwork, _enumerate_templates
    net = GripperRegionNetwork(training=True, group_num=256, gripper_num=64, grasp_score_threshold=0.5, radius=0.06, reg_channel=10)
    assert sum(p.numel() for p in net.parameters()) == 1524396
    keys = set(net.state_dict().keys())
    for k in ("extrat_feature_region.conv.weight", "extrat_feature_region.linear_cls.weight",
              "extrat_feature_region.conv_reg4.bias", "extrat_feature_region.bn_cls4.running_var",
              "extrat_feature_refine.conv_formal.weight", "extrat_feature_refine.bn_formal_reg3.num_batches_tracked"):
        assert k in keys
    assert tuple(net.state_dict()["extrat_feature_refine.conv_formal.weight"].shape) == (1024, 384, 1)
    t = _enumerate_templates()
    assert t.dtype == torch.float16 and tuple(t.shape) == (1, 4, 1, 4) and net.anchor_number == 4
    assert abs(float(t[0, 0, 0, 0]) - 0.57735) < 1e-3 and float(t[0, 0, 0, 0]) != 3 ** 0.5 / 3      # fp16-rounded
    code = ("import sys; sys.path.insert(0, %r); import multi_model.gripper_region_network as m; "
            "print(m.GripperRegionNetwork.__module__)" % os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert out.returncode == 0 and "regnet_for_3d_grasping_b200.gripper_region_network" in out.stdout, out.stderr


def test_region_losses_match_reference_training_branches(monkeypatch):
    """compute_loss (with a ground truth) and compute_loss_refine (with next_gt) of the GripperRegionNetwork mirror against
    the fixture produced by the reference's own methods (oracle/gen_golden_cpu.py region_losses): every element of both
    loss tuples, the confusion counts, the decoded / selected grasps and the masks; np.random.choice is replaced by the
    fixture's fixed rule on both sides."""
    import numpy as np
    from conftest import golden
    from oracle.gen_golden_cpu import deterministic_choice
    from regnet_for_3d_grasping_b200 import gripper_region_network as grn
    ref = golden("ref_py_region_losses.npz")
    monkeypatch.setattr(np.random, "choice", deterministic_choice)
    net = grn.GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                                   reg_channel=10).eval()
    t = lambda k: torch.from_numpy(ref[k])
    anchors = net._enumerate_anchors(t("centers"))
    next_grasp, loss_tuple, correct, next_gt, tt_gt, gmask = net.compute_loss(t("first_grasp"), anchors, t("first_cls"), t("ground"))
    assert torch.equal(gmask, t("l_gmask")) and len(gmask) == 78
    assert torch.allclose(next_grasp, t("l_next_grasp"), rtol=1e-6, atol=1e-7)
    assert torch.allclose(next_gt, t("l_next_gt")) and torch.allclose(tt_gt, t("l_tt_gt"))
    np.testing.assert_allclose([float(x) for x in loss_tuple], ref["l_loss"], rtol=2e-6)
    assert [float(x) for x in correct] == ref["l_correct"].tolist()
    out = net.compute_loss_refine(t("next_grasp"), t("next_x_cls"), t("next_x_reg"), t("next_gt"))
    sel_class, sel_score, sel_stage2, class_select, score_select, loss_refine, correct_refine = out
    assert torch.equal(class_select, t("r_class_select")) and torch.equal(score_select, t("r_score_select"))
    assert torch.allclose(sel_class, t("r_sel_class")) and torch.allclose(sel_score, t("r_sel_score"))
    assert torch.allclose(sel_stage2, t("r_sel_stage2"))
    np.testing.assert_allclose([float(x) for x in loss_refine], ref["r_loss"], rtol=2e-6)
    assert [float(x) for x in correct_refine] == ref["r_correct"].tolist()
    # the stage-1 loss back-propagates into the regressions and the anchor scores
    fg, fc = t("first_grasp").requires_grad_(True), t("first_cls").requires_grad_(True)
    net.compute_loss(fg, anchors, fc, t("ground"))[1][0].backward()
    assert fg.grad.abs().sum() > 0 and fc.grad.abs().sum() > 0
    # inference call unchanged
    ng, lt, ct, gt, tt, gm = net.compute_loss(t("first_grasp"), anchors, t("first_cls"), None)
    assert lt == (None, None) and gt is None and len(gm) == 80


def test_center_grasp_label_lookup_matches_reference(tmp_path):
    """region.get_center_grasp (+ transform_grasp) against the fixture produced by the reference's own _get_center_grasp
    on two synthetic scene files, which the test re-creates from the same seeds: labelled and unlabelled centres, the
    (centre, axis, angle, scores) form and the raw-frame form."""
    from conftest import golden
    from regnet_for_3d_grasping_b200 import region, synth
    ref = golden("ref_py_center_grasp.npz")
    pc = ref["pc"]
    paths = [synth.write_scene_file(str(tmp_path / f"scene{b}.p"), 70 + b, pc[b], n_grasps=10, hit_frac=0.5) for b in range(2)]
    cidx, cpc = torch.from_numpy(ref["center_pc_index"]), torch.from_numpy(ref["center_pc"])
    labels = region.get_center_grasp(cidx, cpc, paths, 0.06, True)
    want = torch.from_numpy(ref["labels"])
    assert tuple(labels.shape) == tuple(want.shape) == (2, 24, 10)
    assert torch.equal(labels[:, :, 7] == -1, want[:, :, 7] == -1) and 0 < int((want[:, :, 7] == -1).sum()) < 48
    assert torch.allclose(labels, want, rtol=1e-6, atol=1e-7)
    frames = region.get_center_grasp(cidx, cpc, paths, 0.06, False)
    assert torch.allclose(frames, torch.from_numpy(ref["frames"]), rtol=1e-6, atol=1e-7)
    # the annotation tensors are cached per (path, mtime, size): same answer from the cache, and a rewritten file is re-read
    assert len(region._SCENE_CACHE) >= 2 and torch.equal(region.get_center_grasp(cidx, cpc, paths, 0.06, True), labels)
    synth.write_scene_file(paths[0], 999, pc[0], n_grasps=11, hit_frac=0.5)
    assert not torch.equal(region.get_center_grasp(cidx, cpc, paths, 0.06, True)[0], labels[0])


def test_import_shims_and_synthetic_dataset(tmp_path):
    """the stand-ins for `tensorboardX` / `transforms3d` that the reference's train.py / utils.py
    import (checked against scipy), and the synthetic training set in the reference's on-disk format -- read back the way
    dataset_utils/scoredataset.py:61-82 reads it (and through the reference's own ScoreDataset where /root/reference
    exists)."""
    import importlib
    import os
    import sys
    import numpy as np
    from scipy.spatial.transform import Rotation
    from regnet_for_3d_grasping_b200 import synth
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin")
    sys.path.insert(0, dropin)
    try:
        t3d = importlib.import_module("transforms3d")
        tbx = importlib.import_module("tensorboardX")
        for a in [(-0.87 * np.pi, 0.0, 0.0), (0.3, -1.1, 2.0), (3.0, 0.2, -0.4)]:
            m = t3d.quaternions.quat2mat(t3d.euler.euler2quat(*a))
            assert np.abs(m - Rotation.from_euler("xyz", a).as_matrix()).max() < 1e-12
        q = t3d.quaternions.axangle2quat([1, 0, 0], np.pi * 1.13)
        assert np.abs(t3d.quaternions.quat2mat(q) - Rotation.from_rotvec([np.pi * 1.13, 0, 0]).as_matrix()).max() < 1e-12
        w = tbx.SummaryWriter(str(tmp_path / "log"))
        w.add_scalar("loss", 1.5, 0)
        w.close()
    finally:
        sys.path.remove(dropin)
    paths = synth.write_dataset(str(tmp_path / "data"), n_scenes=5, seed=3, n_view=4000, n_grasps=50)
    assert len(paths) == 5 and sorted(os.listdir(tmp_path / "data" / "training_data")) == [os.path.basename(p) for p in paths]
    data = np.load(paths[2], allow_pickle=True)
    n = len(data["view_cloud"])
    assert data["view_cloud"].shape == (n, 3) and data["view_cloud_color"].shape == (n, 3)
    assert data["view_cloud_score"].shape == (n,) and data["view_cloud_label"].shape == (n,)
    assert data["frame"].shape == (50, 4, 4) and data["antipodal_score"].shape == (50,) and data["scene_cloud"].shape == (n, 3)
    assert (data["view_cloud_label"] == 0).any() and (data["view_cloud_label"] > 0).any()
    if os.path.isdir("/root/reference/dataset_utils"):      # build container only: the reference's own dataset class
        sys.path.insert(0, "/root/reference")
        try:
            from dataset_utils.scoredataset import ScoreDataset
            ds = ScoreDataset(2048, str(tmp_path / "data"), "train", 1, [0.08])
            view, view_score, view_label, data_path, width = ds[0]
            assert view.shape == (2048, 6) and view_score.shape == (2048,) and len(ds) == 4 and os.path.exists(data_path)
        finally:
            sys.path.remove("/root/reference")


def test_open3d_standin_covers_the_reference_call_sites(tmp_path):
    """The `open3d` stand-in (dropin/open3d): point-cloud container, normals (PCA, oriented to the camera), kd-tree
    searches against brute force, voxel down-sampling, PCD round trips (binary and ascii) -- the calls of test.py:102-106
    and dataset_utils/eval_score/eval_utils/{pointcloud,torch_scene_point_cloud,evaluation_data_generator}.py."""
    import importlib
    import os
    import sys
    import numpy as np
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin")
    sys.path.insert(0, dropin)
    try:
        o3d = importlib.import_module("open3d")
        rng = np.random.default_rng(0)
        pts = rng.random((600, 3))
        pts[:, 2] = 0.002 * rng.random(600)                                   # a thin slab: normals are +-z
        cloud = o3d.geometry.PointCloud()
        cloud.points = o3d.utility.Vector3dVector(pts)
        cloud.colors = o3d.utility.Vector3dVector(rng.random((600, 3)))
        cloud.estimate_normals(search_param=o3d.geometry.KDTreeSearchParamHybrid(radius=0.2, max_nn=30),
                               fast_normal_computation=False)
        cloud.normalize_normals()
        cloud.orient_normals_towards_camera_location(np.array([0.0, 0.0, 5.0]))
        n = np.asarray(cloud.normals)
        assert n.shape == (600, 3) and n[:, 2].min() > 0.99 and np.allclose(np.linalg.norm(n, axis=1), 1.0)
        tree = o3d.geometry.KDTreeFlann(cloud)
        q = pts[7]
        d2 = ((pts - q) ** 2).sum(1)
        k, idx, dist = tree.search_radius_vector_3d(q, 0.15)
        assert k == int((d2 <= 0.15 ** 2).sum()) and sorted(idx) == sorted(np.nonzero(d2 <= 0.15 ** 2)[0].tolist()) and idx[0] == 7
        k, idx, dist = tree.search_knn_vector_3d(q, 5)
        assert k == 5 and idx == np.argsort(d2, kind="stable")[:5].tolist() and np.allclose(dist, np.sort(d2)[:5])
        assert tree.search_hybrid_vector_3d(q, 0.15, 3)[0] == 3
        moved = o3d.geometry.PointCloud()
        moved.points = pts.copy()
        T = np.eye(4)
        T[:3, 3] = [1.0, 2.0, 3.0]
        assert np.allclose(np.asarray(moved.transform(T).points), pts + [1.0, 2.0, 3.0])
        assert 0 < len(cloud.voxel_down_sample(0.25).points) <= 16 + 4
        for ascii_ in (False, True):
            path = str(tmp_path / f"cloud_{int(ascii_)}.pcd")
            o3d.io.write_point_cloud(path, cloud, write_ascii=ascii_)
            back = o3d.io.read_point_cloud(path)
            assert np.abs(np.asarray(back.points) - pts.astype(np.float32)).max() < 1e-7
            assert np.abs(np.asarray(back.colors) - np.asarray(cloud.colors)).max() <= 0.5 / 255 + 1e-9
    finally:
        sys.path.remove(dropin)
        for name in [m for m in sys.modules if m == "open3d" or m.startswith("open3d.")]:
            del sys.modules[name]


def test_region_loss_edge_cases():
    """compute_loss_refine when no stage-1 grasp matches its ground truth (no class-balanced pair, :260) and when nothing
    is classified positive (:274): zero losses / diagnostics, counts still reported -- the reference's behaviour."""
    from regnet_for_3d_grasping_b200 import gripper_region_network as grn
    net = grn.GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                                   reg_channel=10).eval()
    g = torch.Generator().manual_seed(1)
    m = 9
    next_grasp = torch.rand(m, 10, generator=g)
    next_gt = next_grasp.clone()
    next_gt[:, :3] += 1.0                                    # every ground truth 1.7 m away: all of class 0
    cls = torch.stack([torch.ones(m), -torch.ones(m)], dim=1)  # everything predicted negative
    out = net.compute_loss_refine(next_grasp, cls, torch.zeros(m, 10), next_gt)
    sel_class, sel_score, sel_stage2, class_select, score_select, loss_tuple, correct = out
    assert len(class_select) == 0 and len(score_select) == 0 and sel_class.shape == (0, 10)
    assert len(loss_tuple) == 18 and all(float(x) == 0.0 for x in loss_tuple)
    assert [float(x) for x in correct] == [0.0, float(m), 0.0, 0.0]           # TP, TN, FP, FN
    # everything predicted positive, still no positive ground truth: diagnostics are computed, the loss stays 0
    out = net.compute_loss_refine(next_grasp, -cls, torch.zeros(m, 10), next_gt)
    assert len(out[3]) == m and float(out[5][0]) == 0.0 and float(out[5][10]) > 0 and [float(x) for x in out[6]] == [0.0, 0.0, float(m), 0.0]
    # stage 1: a ground truth in which only one anchor orientation occurs (the other anchors have no member)
    centers = torch.rand(6, 3, generator=g)
    anchors = net._enumerate_anchors(centers)
    ground = torch.cat([centers, anchors[:, 0, 3:6], torch.zeros(6, 1), torch.rand(6, 3, generator=g)], dim=1).view(1, 6, 10)
    ng, lt, ct, gt, tt, gm = net.compute_loss(torch.zeros(6, 4, 10), anchors, torch.randn(6, 4, generator=g), ground)
    assert len(gm) == 6 and torch.isfinite(lt[0]) and torch.allclose(tt, anchors[:, 0]) and float(ct[0] + ct[1]) == 6.0


def test_eval_test_view_collision_filter_matches_reference():
    """grasp_eval.eval_test (batched) against the fixture produced by the reference's own EvalDataTest.run_collision_view:
    the same grasps survive, in the same order, and the decoded gripper frames agree."""
    from conftest import golden
    from oracle.gen_golden_cpu import eval_test_inputs
    from regnet_for_3d_grasping_b200 import grasp_eval
    ref = golden("ref_py_eval_test.npz")
    pts, grasp, table_height, depth, width = eval_test_inputs()
    frame, center, score = grasp_eval.grasp_frames(grasp)
    assert torch.allclose(frame, torch.from_numpy(ref["frame"]), atol=1e-6)
    mask = grasp_eval.view_collision_free(pts, grasp, table_height, depth, width, batch=64)
    assert torch.nonzero(mask).view(-1).tolist() == ref["kept_index"].tolist()
    kept = grasp_eval.eval_test(pts.numpy(), grasp, None, table_height, depth, width, -1)
    assert torch.equal(kept, torch.from_numpy(ref["kept"])) and 10 < len(kept) < 390
    assert grasp_eval.eval_test(pts, grasp[:0], None, table_height, depth, width, -1).shape == (0, 8)


def test_eval_validate_matches_reference():
    """grasp_eval.eval_validate (batched view filter, scene filter and antipodal score) against the fixture produced by
    the reference's own EvalDataValidate.run_collision: same grasp sets, same counts, same total score."""
    from conftest import golden
    from oracle.gen_golden_cpu import eval_validate_inputs
    from regnet_for_3d_grasping_b200 import grasp_eval
    ref = golden("ref_py_eval_validate.npz")
    data, grasp, table_height, depth, width = eval_validate_inputs()
    vgr, score, n_view, g_view, g_scene = grasp_eval.eval_validate(data, grasp, 0, table_height, depth, width, -1)
    assert n_view == int(ref["n_view"]) == 300 and vgr == int(ref["vgr"]) == 20
    assert torch.equal(g_view, torch.from_numpy(ref["grasp_view"])) and torch.equal(g_scene, torch.from_numpy(ref["grasp_scene"]))
    assert abs(score - float(ref["score"])) <= 1e-5 * float(ref["score"])
    # one depth per grasp (utils.py passes a tensor when the model predicts the depth): same answer for a constant tensor
    out = grasp_eval.eval_validate(data, grasp, 0, table_height, torch.full((len(grasp),), depth), width, -1)
    assert out[0] == vgr and out[2] == n_view and abs(out[1] - score) < 1e-6


def test_grasp_label_encoding_round_trips_through_the_decoder():
    """Size-independent property tying the two ends of the region stage together: a gripper frame encoded as a label
    (region.transform_grasp: centre, closing axis with x >= 0, angle) and decoded again (grasp_eval.grasp_frames, the same
    algebra as closing_box_frame) gives back the approach axis, the closing axis up to the encoder's sign convention, and
    the minor normal with that sign -- for arbitrary rotations."""
    from scipy.spatial.transform import Rotation
    from regnet_for_3d_grasping_b200 import grasp_eval, region
    from regnet_for_3d_grasping_b200.gripper_region_network import closing_box_frame
    M = 3000
    rot = torch.from_numpy(Rotation.random(M, random_state=0).as_matrix()).float()
    centre = torch.rand(M, 3, generator=torch.Generator().manual_seed(0))
    score = torch.rand(1, M, generator=torch.Generator().manual_seed(1))
    label = region.transform_grasp(torch.cat([rot, centre.view(M, 3, 1)], dim=2).view(1, M, 3, 4), score, score, score)[0]
    assert (label[:, 3] >= 0).all() and (label[:, 6].abs() <= math.pi + 1e-6).all() and torch.equal(label[:, :3], centre)
    frame, c2, s2 = grasp_eval.grasp_frames(label[:, :8])
    sign = torch.where(rot[:, 0, 1] < 0, -1.0, 1.0).view(M, 1)
    assert torch.allclose(frame[:, :, 0], rot[:, :, 0], atol=2e-6)               # approach
    assert torch.allclose(frame[:, :, 1], sign * rot[:, :, 1], atol=2e-6)        # closing axis, flipped to x >= 0
    assert torch.allclose(frame[:, :, 2], sign * rot[:, :, 2], atol=2e-6)        # minor normal follows
    assert torch.equal(c2, centre) and torch.equal(s2.view(-1), score.view(-1))
    # the closing-box crop of the refine stage uses the same frame (rows instead of columns)
    assert torch.allclose(closing_box_frame(label[:, :8]), frame.transpose(1, 2), atol=2e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference checkout (build container only)")
def test_reference_train_py_runs_unchanged_on_the_dropin():
    """The acceptance criterion of the drop-in, as far as it can be exercised without a GPU: the reference's own
    train.py --mode pretrain_score (file untouched, run in a subprocess by tests/dryrun_reference_train.py) completes an
    epoch of training + validation on a synthetic data set with every `multi_model.*` / `pn2_ext` / third-party import
    resolving to this repository, and saves a model whose class is this repository's ScoreNetwork."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "dryrun_reference_train.py")], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "dry run ok: pretrain_score" in r.stdout and "regnet_for_3d_grasping_b200.score_network.ScoreNetwork" in r.stdout
    assert r.stdout.count("train Epoch: 0") == 2 and "validate Epoch: 0" in r.stdout


def test_region_loss_without_any_labelled_centre():
    """a training batch whose centres have no ground-truth grasp must not crash: NaN losses (what the
    reference's empty means give), no grasp handed to the refine stage."""
    import torch
    from regnet_for_3d_grasping_b200.gripper_region_network import GripperRegionNetwork
    net = GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06, reg_channel=10)
    M, A = 6, 4
    first_grasp = torch.randn(M, A, 10, requires_grad=True)
    anchors = torch.randn(M, A, 7)
    first_cls = torch.randn(M, A)
    ground = -torch.ones(2, 3, 10)
    next_grasp, loss_tuple, correct, next_gt, tt_gt, gmask = net.compute_loss(first_grasp, anchors, first_cls, ground)
    assert next_grasp.shape == (0, 10) and next_gt.shape == (0, 10) and tt_gt.shape == (0, 7) and gmask.numel() == 0
    assert len(loss_tuple) == 10 and bool(torch.isnan(loss_tuple[0])) and loss_tuple[0].requires_grad
    assert float(correct[0]) == 0.0 and float(correct[1]) == 0.0


def test_eval_validate_estimates_missing_scene_normals():
    """a scene file without 'scene_normal' gets its normals estimated (as the reference does through open3d)
    instead of raising; the estimate is cached on the dictionary."""
    import numpy as np
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dropin = os.path.join(root, "regnet_for_3d_grasping_b200", "dropin")
    if dropin not in sys.path:
        sys.path.insert(0, dropin)
    from regnet_for_3d_grasping_b200 import grasp_eval
    rng = np.random.default_rng(0)
    xy = rng.uniform(-0.1, 0.1, size=(4000, 2))
    scene = np.c_[xy, 0.75 + 0.001 * rng.standard_normal(4000)].astype(np.float32)       # a table top
    normals = grasp_eval._estimate_scene_normals(scene)
    assert normals.shape == (4000, 3) and np.abs(np.abs(normals[:, 2]) - 1.0).mean() < 0.05
    d = {"view_cloud": scene[::4].copy(), "scene_cloud": scene}
    grasp = torch.tensor([[0.0, 0.0, 0.80, 0.0, 1.0, 0.0, 0.0, 0.5]])
    out = grasp_eval.eval_validate(d, grasp, None, 0.75, 0.06, 0.08, -1)
    assert "_estimated_scene_normal" in d and out is not None


def test_level0_recompute_closed_forms_match_autograd():
    """The algebra behind csrc/train_gather.cu "set-abstraction level 0, first block" (no stored pre-activation), in float64
    on the CPU: batch moments of z = W x from the 6 sums and 21 products of x, and dW / dgamma / dbeta of
    y = relu(bn(z)) from 7 sums per channel over g = dy * [bn(z) > 0] -- against autograd through the plain formulation."""
    torch.manual_seed(3)
    P, C, eps = 4096, 16, 1e-5
    x = torch.randn(P, 6, dtype=torch.float64) * torch.tensor([0.02, 0.02, 0.02, 0.5, 0.5, 0.5], dtype=torch.float64) + 0.1
    W = torch.randn(C, 6, dtype=torch.float64, requires_grad=True)
    gamma = (torch.rand(C, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = (torch.rand(C, dtype=torch.float64) - 0.5).requires_grad_(True)
    z = x @ W.t()
    mu, var = z.mean(0), z.var(0, unbiased=False)
    y = torch.relu((z - mu) / torch.sqrt(var + eps) * gamma + beta)
    dy = torch.randn(P, C, dtype=torch.float64)
    (y * dy).sum().backward()
    with torch.no_grad():
        sx, sxx = x.sum(0), x.t() @ x                              # what sa0_input_moments_kernel accumulates (27 numbers)
        m1, m2 = W @ sx, ((W @ sxx) * W).sum(1)                    # sa0_moments_from_sums_kernel
        assert torch.allclose(m1 / P, mu, rtol=1e-12, atol=1e-14)
        assert torch.allclose(m2 / P - (m1 / P) ** 2, var, rtol=1e-9, atol=1e-14)
        istd = 1.0 / torch.sqrt(var + eps)
        sc = gamma * istd
        g = dy * (((z - mu) * sc + beta) > 0)                      # sa0_backward_sums_kernel: mask recomputed from the inputs
        g0, gx = g.sum(0), g.t() @ x                               # the 7 sums per channel
        dbeta = g0                                                 # sa0_backward_finalize_kernel
        dgamma = istd * ((W * gx).sum(1) - mu * g0)
        zx = W @ sxx - mu[:, None] * sx[None, :]
        dW = sc[:, None] * (gx - (g0 / P)[:, None] * sx[None, :] - (dgamma * istd / P)[:, None] * zx)
    assert torch.allclose(dbeta, beta.grad, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dgamma, gamma.grad, rtol=1e-9, atol=1e-10)
    assert torch.allclose(dW, W.grad, rtol=1e-8, atol=1e-9)


def test_folded_set_abstraction_operand_is_the_grouped_first_block():
    """The fold behind csrc/gemm_fused_a.cu, in float64 on the CPU: relu(scale (W_f f_j + W_x (xyz_j - c_m)) + shift) ==
    relu(Z'[j] + T[m]) with Z' = scale (W_f f + W_x xyz) per source point and T = shift - scale W_x c per centroid."""
    torch.manual_seed(4)
    N, M, K, Cf, C0 = 200, 20, 8, 12, 10
    f = torch.randn(N, Cf, dtype=torch.float64)
    xyz = torch.rand(N, 3, dtype=torch.float64)
    ctr = xyz[torch.randperm(N)[:M]]
    nbr = torch.randint(0, N, (M, K))
    Wf, Wx = torch.randn(C0, Cf, dtype=torch.float64), torch.randn(C0, 3, dtype=torch.float64)
    scale, shift = torch.rand(C0, dtype=torch.float64) + 0.5, torch.randn(C0, dtype=torch.float64)
    grouped = torch.relu((f[nbr] @ Wf.t() + (xyz[nbr] - ctr[:, None, :]) @ Wx.t()) * scale + shift)      # (M, K, C0)
    zp = (f @ Wf.t() + xyz @ Wx.t()) * scale
    t = shift - (ctr @ Wx.t()) * scale
    folded = torch.relu(zp[nbr] + t[:, None, :])
    assert torch.allclose(folded, grouped, rtol=1e-12, atol=1e-12)
