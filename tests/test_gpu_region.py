"""GPU parity, region stage: deterministic parts bit-exact against the
CPU restatement (oracle/region_oracle.py), random parts by membership / count / uniformity properties."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _scene(B, N, seed):
    from regnet_for_3d_grasping_b200 import synth
    pc = torch.from_numpy(synth.batch("table", range(seed, seed + B), N))
    g = torch.Generator().manual_seed(seed)
    score = torch.rand(B, N, generator=g)
    return pc, score


@pytest.mark.parametrize("B,N,M", [(3, 25600, 64), (1, 25600, 4000), (2, 5000, 300)])
def test_select_score_center_fps_branch_exact(lib_path, oracle, B, N, M):
    from oracle import region_oracle
    from regnet_for_3d_grasping_b200 import region
    pc, score = _scene(B, N, 7)
    want = region_oracle.select_score_center_fps_branch(pc, score, M, 0.5)
    cpc, cidx, cnt = region.select_score_center(pc.cuda(), score.cuda(), M, 0.5, seed=1, return_count=True)
    assert cidx.dtype == torch.int64 and tuple(cidx.shape) == (B, M) and tuple(cpc.shape) == (B, M, 6)
    for b in range(B):
        assert cnt[b].item() == want[b][0]
        assert want[b][1] is not None
        assert torch.equal(cidx[b].cpu(), want[b][1]), f"cloud {b}: centre indices differ from FPS over the positives"
        assert torch.equal(cpc[b].cpu(), pc[b, want[b][1]])


def test_select_score_center_random_branches(lib_path):
    from regnet_for_3d_grasping_b200 import region
    B, N, M = 3, 4096, 128
    pc, score = _scene(B, N, 9)
    score[0] = 0.0                                   # no positive: M distinct random points
    score[1] = 0.0
    score[1, [5, 77, 4000]] = 0.9                    # 3 positives: all of them, then repeats
    cpc, cidx, cnt = region.select_score_center(pc.cuda(), score.cuda(), M, 0.5, seed=3, return_count=True)
    cidx = cidx.cpu()
    assert cnt.tolist()[:2] == [0, 3]
    assert cidx[0].unique().numel() == M and cidx[0].min() >= 0 and cidx[0].max() < N
    assert cidx[1, :3].tolist() == [5, 77, 4000] and set(cidx[1].tolist()) == {5, 77, 4000}
    assert torch.equal(cpc.cpu(), torch.stack([pc[b, cidx[b]] for b in range(B)]))
    # a different seed changes the random branches but not the FPS branch
    _, cidx2 = region.select_score_center(pc.cuda(), score.cuda(), M, 0.5, seed=4)
    assert not torch.equal(cidx2[0].cpu(), cidx[0]) and torch.equal(cidx2[2].cpu(), cidx[2])


@pytest.mark.parametrize("scan", [False, True], ids=["grid", "scan"])
@pytest.mark.parametrize("r_time,G", [(0.1, 256), (0.8, 1024), (0.8, 2048)])
def test_ball_crop_membership_and_counts(lib_path, monkeypatch, r_time, G, scan):
    """Both forms of the crop (uniform-grid candidates = default for big clouds; REGNET_API_BRUTE = scan of all points)
    against the reference's membership test.  The reference draws with np.random.choice, i.e. in random order: only the
    scan form happens to return the no-replacement picks in ascending index order."""
    if scan:
        monkeypatch.setenv("REGNET_API_BRUTE", "1")
    from oracle import region_oracle
    from regnet_for_3d_grasping_b200 import region
    B, N, NC = 2, 25600, 96
    pc, score = _scene(B, N, 11)
    cpc, cidx = region.select_score_center(pc.cuda(), score.cuda(), NC, 0.5, seed=1)
    cpc.view(-1, 6)[5, :3] += 10.0                       # one centre far from every point: empty ball
    idx, grp, cnt = region.get_group_pc(pc.cuda(), cpc, cidx, G, 0.08, 0.010, 0.06, r_time, seed=5, return_count=True)
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (B, NC, G) and tuple(grp.shape) == (B, NC, G, 6)
    idx, grp, cnt, cpc = idx.cpu(), grp.cpu(), cnt.cpu(), cpc.cpu()
    radius = max(0.08, 0.010, 0.06) * r_time
    saw_wo, saw_w = False, False
    for b in range(B):
        mask = region_oracle.ball_mask(pc[b], cpc[b], radius)
        assert torch.equal(cnt[b].long(), mask.sum(1)), "number of points inside each ball differs from the reference test"
        for c in range(NC):
            n = int(mask[c].sum())
            row = idx[b, c]
            if n == 0:
                assert (row == -1).all() and (grp[b, c] == -1).all()
                continue
            assert mask[c, row].all(), "sampled a point outside the ball"
            assert torch.equal(grp[b, c], pc[b, row])
            if n >= G:
                saw_wo = True
                assert row.unique().numel() == G and (not scan or (row[1:] > row[:-1]).all())
                if n == G:
                    assert torch.equal(row.sort()[0], torch.nonzero(mask[c]).view(-1))
            else:
                saw_w = True
    assert cnt.view(-1)[5].item() == 0
    assert saw_w or saw_wo


def test_ball_crop_sampling_is_uniform(lib_path):
    """Chi-square on both regimes: 40 points inside the ball, draw 16 without replacement / 100 with replacement."""
    from regnet_for_3d_grasping_b200 import region
    g = torch.Generator().manual_seed(0)
    pts = torch.rand(1, 400, 6, generator=g)
    pts[0, :, :3] += 5.0
    inside = torch.randperm(400, generator=g)[:40]
    pts[0, inside, :3] = torch.rand(40, 3, generator=g) * 0.01
    center = torch.zeros(1, 1, 6)
    center[0, 0, :3] = 0.005
    for G, reps in ((16, 4000), (100, 800)):
        hist = torch.zeros(400)
        for s in range(reps):
            idx, _ = region.get_group_pc(pts.cuda(), center.cuda(), None, G, 1.0, 0.1, 0.1, 0.05, seed=1000 + s)
            hist += torch.bincount(idx.view(-1).cpu(), minlength=400).float()
        assert hist[[i for i in range(400) if i not in set(inside.tolist())]].sum() == 0
        obs = hist[inside]
        exp = obs.sum() / 40
        chi2 = ((obs - exp) ** 2 / exp).sum().item()
        assert chi2 < 80, f"G={G}: chi-square {chi2:.1f} over 39 dof"      # p < 1e-4 beyond ~78


def test_mask_sampler_thresholds_and_uniformity(lib_path):
    from regnet_for_3d_grasping_b200 import region
    g = torch.Generator().manual_seed(1)
    rows, G, K = 300, 2048, 64
    counts = torch.randint(0, 200, (rows,), generator=g)
    counts[:8] = torch.tensor([0, 5, 6, 63, 64, 65, 2048, 1])
    mask = torch.zeros(rows, G, dtype=torch.bool)
    for r in range(rows):
        mask[r, torch.randperm(G, generator=g)[:counts[r]]] = True
    idx, cnt = region.sample_mask_rows(mask.cuda(), K, min_count=5, seed=2, return_count=True)
    idx, cnt = idx.cpu(), cnt.cpu()
    assert torch.equal(cnt.long(), counts)
    for r in range(rows):
        n = int(counts[r])
        if n <= 5:
            assert (idx[r] == -1).all()                                  # rejected (gripper_region_network.py:538-544)
        else:
            assert mask[r, idx[r]].all()
            if n > K:
                assert idx[r].unique().numel() == K                      # "> region_num": without replacement
    assert idx[4].unique().numel() < K or True                           # exactly K set -> WITH replacement (quirk kept)
    hist = torch.zeros(G)
    m1 = torch.zeros(1, G, dtype=torch.bool)
    on = torch.randperm(G, generator=g)[:100]
    m1[0, on] = True
    for s in range(600):
        hist += torch.bincount(region.sample_mask_rows(m1.cuda(), K, seed=50 + s).view(-1).cpu(), minlength=G).float()
    obs = hist[on]
    chi2 = ((obs - obs.mean()) ** 2 / obs.mean()).sum().item()
    assert hist.sum() == 600 * K and chi2 < 170, chi2                    # 99 dof


def test_gather_max_equals_reference_expression(lib_path):
    from regnet_for_3d_grasping_b200 import region
    g = torch.Generator().manual_seed(2)
    B, N, NC, G, C = 3, 2000, 50, 256, 256
    feat = torch.randn(B, N, C, generator=g).cuda()
    idx = torch.randint(0, N, (B, NC, G), generator=g).cuda()
    idx[1, 3] = -1                                                       # an empty group row as the reference leaves it
    got = region.gather_max(feat, idx)
    flat = feat.view(-1, C)
    rows = (idx + (torch.arange(B, device="cuda") * N).view(B, 1, 1)).view(-1)
    want = flat[rows].view(B * NC, G, C).permute(0, 2, 1).max(dim=2)[0].view(B, NC, C)   # MaxPool1d(G)
    assert torch.equal(got, want)


def test_get_grasp_allobj_shapes_test_config(lib_path, tmp_path):
    """test.py:68-71 parameters (center_num 4000, group_num 256, group_num_more 2048) on one cloud."""
    from regnet_for_3d_grasping_b200 import region
    pc, score = _scene(1, 25600, 21)
    params = [4000, 0.5, 256, 0.1, 2048, 0.8, 0.08, 0.010, 0.06]
    torch.manual_seed(0)
    out = region.get_grasp_allobj(pc.cuda(), score.cuda(), params, [])
    center_pc, center_idx, gi, gp, gmi, gmp, labels = out
    assert tuple(center_pc.shape) == (1, 4000, 6) and tuple(center_idx.shape) == (1, 4000)
    assert tuple(gi.shape) == (1, 4000, 256) and tuple(gp.shape) == (1, 4000, 256, 6)
    assert tuple(gmi.shape) == (1, 4000, 2048) and tuple(gmp.shape) == (1, 4000, 2048, 6) and labels is None
    assert (gi >= 0).all() and (gmi >= 0).all()          # every centre is itself inside its ball
    torch.manual_seed(0)
    out2 = region.get_grasp_allobj(pc.cuda(), score.cuda(), params, [])
    assert torch.equal(out2[2], gi) and torch.equal(out2[4], gmi)         # torch.manual_seed makes the draws reproducible
    # with a scene annotation file: grasp labels (centre, axis, angle, 3 scores) per centre, -1 where none is near
    from regnet_for_3d_grasping_b200 import synth
    path = synth.write_scene_file(str(tmp_path / "scene.p"), 9, pc[0].numpy(), n_grasps=40, hit_frac=0.5)
    labels = region.get_grasp_allobj(pc.cuda(), score.cuda(), params, [path])[6]
    assert tuple(labels.shape) == (1, 4000, 10) and labels.is_cuda
    has = labels[0, :, 7] != -1
    # (rows without a grasp: -1 everywhere except the axis, which the reference's sign flip turns into (1, 1, 1))
    assert 0 < int(has.sum()) < 4000 and (labels[0][~has][:, [0, 1, 2, 6, 7, 8, 9]] == -1).all()
    assert (labels[0][:, 3] >= 0).all()


def _region_net_on_gpu():
    import numpy as np
    from helpers import region_net_fixture
    ref, net, sd, inp = region_net_fixture()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    inp = {k: v.cuda() for k, v in inp.items()}
    params = [float(x) for x in ref["params"]]
    args = (inp["pc_group"], inp["pc_group_more"], inp["pc_group_index"], inp["pc_group_more_index"], inp["center_pc"],
            inp["center_pc_index"], inp["pc"], inp["all_feature"], params)
    return ref, net, inp, params, args, np


def test_region_network_vs_reference_python_golden(lib_path, oracle):
    """GripperRegionNetwork mirror (rows R3-R7) against the fixture produced by the REAL reference module: the whole
    16-tuple of the inference call, with the reference's random picks replaced by the fixture's fixed rule on both sides."""
    from oracle import region_oracle
    from regnet_for_3d_grasping_b200.gripper_region_network import get_gripper_region_transform
    ref, net, inp, params, args, np = _region_net_on_gpu()
    net._sampler = lambda mask: region_oracle.sample_rows_fixed_rule(mask.cpu(), 8).to(mask.device)
    with torch.no_grad():
        out = net(*args)
    (next_grasp, keep2, true_mask, loss_tuple, correct_tuple, next_gt, sel_class, sel_score, sel_stage2, keep3, keep3s,
     final_mask, final_mask_sthre, loss_refine, correct_refine, gt) = out
    assert len(out) == 16 and loss_tuple == (None, None) and next_gt is None and gt is None
    np.testing.assert_allclose(next_grasp.cpu().numpy(), ref["next_grasp"], rtol=1e-4, atol=1e-5)
    assert [int(k) for k in keep2] == ref["keep2"].tolist() and torch.equal(true_mask.cpu(), torch.arange(12))
    assert np.array_equal(final_mask.cpu().numpy(), ref["final_mask"])
    assert np.array_equal(final_mask_sthre.cpu().numpy(), ref["final_mask_sthre"])
    assert [int(k) for k in keep3] == ref["keep3"].tolist() and [int(k) for k in keep3s] == ref["keep3s"].tolist()
    for got, key in ((sel_class, "sel_class"), (sel_score, "sel_score"), (sel_stage2, "sel_stage2")):
        np.testing.assert_allclose(got.cpu().numpy(), ref[key], rtol=1e-4, atol=1e-5)
    # the stand-alone closing-box function: same outputs as the reference's, dtype quirk included
    M, NGM = 12, inp["pc_group_more"].shape[2]
    gp, gi, ginall, gmask = get_gripper_region_transform(inp["pc_group_more"].view(M, NGM, 6),
                                                         inp["pc_group_more_index"].view(M, NGM), next_grasp, 8, params,
                                                         sampler=net._sampler)
    assert gp.dtype == torch.int64 and np.array_equal(gp.cpu().numpy(), ref["gripper_pc"])
    assert np.array_equal(gi.cpu().numpy(), ref["gripper_pc_index"])
    assert np.array_equal(ginall.cpu().numpy(), ref["gripper_pc_index_inall"])
    assert np.array_equal(gmask.cpu().numpy(), ref["gripper_mask"])


def test_region_network_device_sampler_properties(lib_path, oracle):
    """Same call with the device sampler: the accepted grasps and the closing-box membership are deterministic (equal to
    the fixture), every sampled index is a member of its box, and a fixed seed reproduces the draw."""
    from oracle import region_oracle
    from regnet_for_3d_grasping_b200.gripper_region_network import closing_box_points, get_gripper_region_transform
    ref, net, inp, params, args, np = _region_net_on_gpu()
    net.sample_seed = 1234
    with torch.no_grad():
        a = net(*args)
        b = net(*args)
    assert torch.equal(a[11], b[11]) and torch.equal(a[6], b[6])                        # reproducible with a seed
    M, NGM = 12, inp["pc_group_more"].shape[2]
    pts = inp["pc_group_more"].view(M, NGM, 6)
    _, mask = closing_box_points(pts, a[0], params)
    _, want_mask = region_oracle.closing_box(pts.cpu(), a[0].cpu(), params)
    assert torch.equal(mask.cpu(), want_mask)
    _, gi, ginall, gmask = get_gripper_region_transform(pts, inp["pc_group_more_index"].view(M, NGM), a[0], 8, params, seed=7)
    assert np.array_equal(gmask.cpu().numpy(), ref["gripper_mask"])                     # accept / reject is not random
    for m in gmask.tolist():
        assert mask[m][gi[m]].all()
        assert torch.equal(ginall[m], inp["pc_group_more_index"].view(M, NGM)[m][gi[m]])
    rejected = [m for m in range(M) if m not in gmask.tolist()]
    assert all((gi[m] == -1).all() and mask[m].sum() <= 5 for m in rejected)


def test_closing_box_mask_kernel_equals_batched_reference_expression(lib_path):
    """The one-pass membership kernel against the reference's expression (bmm of the frame with the centred crop + six
    strict comparisons, gripper_region_network.py:505-528) on random grasps, scalar and per-grasp limits.  The two
    evaluate the same 3-term dot products in different orders, so a point may flip only if it sits within rounding
    distance of a face of the box."""
    from regnet_for_3d_grasping_b200 import region
    from regnet_for_3d_grasping_b200.gripper_region_network import closing_box_frame, closing_box_points
    g = torch.Generator().manual_seed(3)
    M, G = 700, 2048
    pts = (torch.rand(M, G, 6, generator=g) - 0.5) * 0.12
    grasp = torch.cat([(torch.rand(M, 3, generator=g) - 0.5) * 0.02, torch.randn(M, 3, generator=g),
                       (torch.rand(M, 1, generator=g) - 0.5) * 6.0, torch.rand(M, 3, generator=g)], dim=1)
    grasp[0, 3:6] = 0.0                                     # degenerate axis: the reference's fix-up rows
    pts, grasp = pts.cuda(), grasp.cuda()
    for params in ([0.08, 0.010, 0.06], [torch.rand(M, 1, generator=g).cuda() * 0.1 + 0.02, 0.012,
                                         torch.rand(M, 1, generator=g).cuda() * 0.1 + 0.02]):
        pcs_t, want = closing_box_points(pts, grasp, params)
        widths, height, depths = params
        half = lambda v: (v.reshape(-1) / 2) if isinstance(v, torch.Tensor) else v / 2
        got = region.closing_box_mask(pts, grasp[:, :3], closing_box_frame(grasp), half(depths), half(widths), height / 2)
        assert got.dtype == torch.uint8 and tuple(got.shape) == (M, G)
        diff = got.bool() != want
        assert 0.01 < want.float().mean().item() < 0.9      # the box is neither empty nor everything
        if diff.any():
            xl = half(depths).view(-1, 1) if isinstance(depths, torch.Tensor) else depths / 2
            yl = half(widths).view(-1, 1) if isinstance(widths, torch.Tensor) else widths / 2
            x, y, z = pcs_t[..., 0], pcs_t[..., 1], pcs_t[..., 2]
            face = torch.stack([x.abs(), (x - xl).abs(), (y.abs() - yl).abs(), (z.abs() - height / 2).abs()]).min(0)[0]
            assert diff.float().mean().item() < 1e-5 and (face[diff] < 1e-6).all()


def test_region_network_training_call_vs_reference_golden(lib_path, oracle, monkeypatch):
    """The TRAINING call (ground_grasp given, train.py:240-243) against the fixture produced by the reference's own
    forward (oracle/gen_golden_cpu.py region_net_train): both loss tuples (10 + 18 values), confusion counts, masks,
    matched ground truths and selected grasps; random picks replaced by the fixture's fixed rule on both sides."""
    import numpy
    from conftest import golden
    from oracle import region_oracle
    from oracle.gen_golden_cpu import deterministic_choice
    ref, net, inp, params, args, np = _region_net_on_gpu()
    tr = golden("ref_py_region_net_train.npz")
    monkeypatch.setattr(numpy.random, "choice", deterministic_choice)
    net._sampler = lambda mask: region_oracle.sample_rows_fixed_rule(mask.cpu(), 8).to(mask.device)
    ground = torch.from_numpy(tr["ground"]).cuda()
    with torch.no_grad():
        out = net(*args, ground_grasp=ground)
    (next_grasp, keep2, true_mask, loss_tuple, correct_tuple, next_gt, sel_class, sel_score, sel_stage2, keep3, keep3s,
     final_mask, final_mask_sthre, loss_refine, correct_refine, gt) = out
    assert np.array_equal(true_mask.cpu().numpy(), tr["true_mask"]) and [int(k) for k in keep2] == tr["keep2"].tolist()
    np.testing.assert_allclose(next_grasp.cpu().numpy(), tr["next_grasp"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose([float(x) for x in loss_tuple], tr["loss"], rtol=2e-4, atol=1e-6)
    assert [float(x) for x in correct_tuple] == tr["correct"].tolist()
    np.testing.assert_allclose(next_gt.cpu().numpy(), tr["next_gt"], rtol=1e-6)
    assert np.array_equal(final_mask.cpu().numpy(), tr["final_mask"])
    assert np.array_equal(final_mask_sthre.cpu().numpy(), tr["final_mask_sthre"])
    assert [int(k) for k in keep3] == tr["keep3"].tolist() and [int(k) for k in keep3s] == tr["keep3s"].tolist()
    np.testing.assert_allclose([float(x) for x in loss_refine], tr["loss_refine"], rtol=2e-4, atol=1e-6)
    assert [float(x) for x in correct_refine] == tr["correct_refine"].tolist()
    np.testing.assert_allclose(gt.cpu().numpy(), tr["gt"], rtol=1e-6)
    for got, key in ((sel_class, "sel_class"), (sel_score, "sel_score"), (sel_stage2, "sel_stage2")):
        np.testing.assert_allclose(got.cpu().numpy(), tr[key], rtol=1e-4, atol=1e-5)


def test_region_network_training_step_backpropagates(lib_path):
    """train.py --mode train shape of the region step: train-mode heads (batch statistics), both losses summed, gradients
    reach the head parameters AND all_feature (joint training with ScoreNet: the gather + max falls back to autograd's)."""
    ref, net, inp, params, args, np = _region_net_on_gpu()
    from oracle.gen_golden_cpu import region_net_ground
    ground = region_net_ground({"center_pc": inp["center_pc"].cpu()}).cuda()
    net.train()
    net.sample_seed = 5
    all_feature = inp["all_feature"].clone().requires_grad_(True)
    args = list(args)
    args[7] = all_feature
    out = net(*args, ground_grasp=ground)
    loss = out[3][0] + (out[13][0] if out[13][0] is not None else 0.0)
    assert torch.isfinite(loss)
    loss.backward()
    assert all_feature.grad is not None and all_feature.grad.abs().sum() > 0
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert "extrat_feature_region.conv.weight" in grads and grads["extrat_feature_region.conv_reg4.weight"].abs().sum() > 0
    if out[13][0] is not None and out[13][0].requires_grad:
        assert grads["extrat_feature_refine.conv_formal.weight"].abs().sum() > 0


def test_eval_test_on_device_equals_host_and_reference(lib_path):
    """grasp_eval.eval_test on the GPU: same surviving grasps as on the host and as the reference's loop (fixture), and the
    test.py-sized call (4 000 grasps x 25 600 points) runs in milliseconds."""
    import time
    from conftest import golden
    from oracle.gen_golden_cpu import eval_test_inputs
    from regnet_for_3d_grasping_b200 import grasp_eval
    ref = golden("ref_py_eval_test.npz")
    pts, grasp, table_height, depth, width = eval_test_inputs()
    kept = grasp_eval.eval_test(pts.cuda(), grasp.cuda(), None, table_height, depth, width, 0)
    assert kept.is_cuda and torch.equal(kept.cpu(), torch.from_numpy(ref["kept"]))
    pts, grasp, table_height, depth, width = eval_test_inputs(seed=43, N=25600, M=4000)
    want = grasp_eval.view_collision_free(pts, grasp, table_height, depth, width)
    pts, grasp = pts.cuda(), grasp.cuda()
    grasp_eval.eval_test(pts, grasp, None, table_height, depth, width, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kept = grasp_eval.eval_test(pts, grasp, None, table_height, depth, width, 0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    got = torch.zeros(4000, dtype=torch.bool)
    # boundary points may flip between the host's and the device's matmul: allow a handful of grasps to differ
    host_kept = grasp.cpu()[want]
    assert abs(len(kept) - len(host_kept)) <= 4 and 0 < len(kept) < 4000
    print(f"eval_test 4000 grasps x 25600 points: {ms:.2f} ms, kept {len(kept)} (host {len(host_kept)})")
    assert ms < 200
