"""CPU suite, part 1: the oracle against every fixture we have for this path.

ref_py_*.npz   : outputs of the REAL reference Python modules (oracle/gen_golden_cpu.py, build container)
ref_cuda_ops.npz: outputs of the REAL reference CUDA kernels on a B200 (oracle/gen_golden_gpu.py)
"""
import numpy as np
import pytest
import torch

from conftest import golden
from helpers import bq_rowhash


def test_oracle_vs_reference_cuda_kernels(oracle):
    """Pins the C restatement: bit-equal indices / counts / squared distances on every seeded case."""
    ref = golden("ref_cuda_ops.npz")
    from oracle import gen_golden_gpu
    for name, c in gen_golden_gpu.cases().items():
        pc = torch.from_numpy(c["pts"])
        xyz = pc[:, :, :3].permute(0, 2, 1)
        B = xyz.shape[0]
        idx = oracle.farthest_point_sample(xyz, c["M"])
        assert np.array_equal(idx.numpy(), ref[name + ".fps"]), f"{name}: FPS differs from the reference kernel"
        new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, c["M"]))
        bq, cnt = oracle.ball_query(xyz, new_xyz, c["radius"], c["K"])
        assert np.array_equal(cnt.numpy(), ref[name + ".bqcnt"]), f"{name}: ball-query count"
        if name + ".bq" in ref.files:
            assert np.array_equal(bq.numpy(), ref[name + ".bq"]), f"{name}: ball-query index"
        else:
            assert np.array_equal(bq.numpy()[:, :256], ref[name + ".bq_head"])
            assert np.array_equal(bq_rowhash(bq.numpy(), c["K"]), ref[name + ".bq_rowhash"])
        if c["M"] >= 3:
            nn, nnd = oracle.point_search(xyz, new_xyz, 3)
            if name + ".nn" in ref.files:
                assert np.array_equal(nn.numpy(), ref[name + ".nn"]), f"{name}: 3-NN index"
                assert np.array_equal(nnd.numpy(), ref[name + ".nnd"]), f"{name}: 3-NN squared distance"
            else:
                assert np.array_equal(nn.numpy()[:, :4096], ref[name + ".nn_head"])
                assert np.array_equal(nnd.numpy()[:, :4096], ref[name + ".nnd_head"])


def test_oracle_float_ops_vs_reference_cuda(oracle):
    ref = golden("ref_cuda_ops.npz")
    from oracle import gen_golden_gpu
    g = torch.Generator().manual_seed(123)
    c = gen_golden_gpu.cases()["cube_1024_b2"]
    pc = torch.from_numpy(c["pts"])
    xyz = pc[:, :, :3].permute(0, 2, 1)
    feat = torch.randn(2, 19, 1024, generator=g)
    idx = oracle.farthest_point_sample(xyz, 256)
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(2, 3, 256))
    bq, _ = oracle.ball_query(xyz, new_xyz, 0.12, 16)
    grouped = oracle.group_points_forward(feat, bq)
    gout = torch.randn(2, 19, 256, 16, generator=g)
    ggrad = oracle.group_points_backward(gout, bq, 1024)
    nn, nnd = oracle.point_search(xyz, new_xyz, 3)
    inv = 1.0 / torch.clamp(nnd, min=1e-10)
    w = inv / inv.sum(2, keepdim=True)
    sfeat = torch.randn(2, 19, 256, generator=g)
    interp = oracle.interpolate_forward(sfeat, nn, w)
    iout = torch.randn(2, 19, 1024, generator=g)
    igrad = oracle.interpolate_backward(iout, nn, w, 256)
    assert np.array_equal(grouped.numpy(), ref["float.grouped"])          # pure gather: exact
    np.testing.assert_allclose(w.numpy(), ref["float.weight"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(interp.numpy(), ref["float.interp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ggrad.numpy(), ref["float.group_grad"], rtol=1e-5, atol=1e-5)   # atomics order
    np.testing.assert_allclose(igrad.numpy(), ref["float.interp_grad"], rtol=1e-5, atol=1e-5)


def test_oracle_ops_vs_reference_python_small(oracle):
    """The ops as the real reference modules saw them (ref_py_modules_small.npz was produced THROUGH the
    reference's function.py / modules.py with this oracle underneath, so this is a regression pin of the C file
    plus a check that fixture and seeds still line up)."""
    ref = golden("ref_py_modules_small.npz")
    pc = torch.from_numpy(ref["pc"])
    xyz = pc[:, :, :3].permute(0, 2, 1)
    idx = oracle.farthest_point_sample(xyz, 128)
    assert np.array_equal(idx.numpy(), ref["fps"])
    new_xyz = torch.from_numpy(ref["new_xyz"])
    bq, cnt = oracle.ball_query(xyz, new_xyz, 0.15, 16)
    assert np.array_equal(bq.numpy(), ref["bq"]) and np.array_equal(cnt.numpy(), ref["bqcnt"])
    nn, nnd = oracle.point_search(xyz, new_xyz, 3)
    assert np.array_equal(nn.numpy(), ref["nn"]) and np.array_equal(nnd.numpy(), ref["nnd"])


def test_restated_scorenet_vs_reference_python(oracle):
    """oracle/ref_modules.py (the restatement that travels to the GPU box) against the fixture produced by the
    reference's own ScoreNetwork: FPS indices exact, features/scores equal to fp32 round-off."""
    ref = golden("ref_py_scorenet_n6144.npz")
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    pc = torch.from_numpy(synth.batch("table", [11], 6144))
    sd = ref_modules.random_scorenet_state(seed=3)
    with torch.no_grad():
        feat, score, dbg = ref_modules.scorenet_forward(sd, pc, oracle.as_pn2_ext(), keep=True)
    for i in range(3):
        assert np.array_equal(dbg[f"fps{i}"][0].numpy(), ref[f"fps{i}"])
        assert np.array_equal(dbg[f"bqcnt{i}"][0].numpy(), ref[f"bqcnt{i}"])
        assert np.array_equal(dbg[f"bq{i}"][0].sum(1).numpy(), ref[f"bq{i}_sum"])
    assert np.array_equal(dbg["nn2"][0, ::13].numpy(), ref["nn2"])
    rows = ref["rows"]
    np.testing.assert_allclose(feat[0, rows].numpy(), ref["all_feature_rows"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(score[0].numpy(), ref["score"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dbg["sa2"][0, :, ::16].numpy(), ref["sa2_rows"], rtol=1e-5, atol=1e-6)
    # loss path (score_network.py:50-51)
    tgt = torch.from_numpy(synth.scores_like_dataset(5, 1, 6144))
    loss = torch.nn.functional.mse_loss(score, tgt)
    np.testing.assert_allclose(loss.numpy(), golden("ref_py_scorenet_loss.npz")["loss"], rtol=1e-5)


@pytest.mark.parametrize("kind,n,m", [("lattice", 700, 300), ("cube", 513, 200), ("lattice", 40, 40), ("cube", 17, 9)])
def test_fps_tie_rule_closed_form(oracle, kind, n, m):
    """Block emulation (C) == closed-form bit-reversed-slot rule (numpy), incl. tie-heavy lattices."""
    from regnet_for_3d_grasping_b200 import synth
    pts = synth.batch(kind, [n + m], n)
    xyz = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1)
    a = oracle.farthest_point_sample(xyz, m)[0].numpy()
    b = oracle.np_fps_tierule(pts[0, :, :3], m)
    assert np.array_equal(a, b)


def test_oracle_edge_cases(oracle):
    xyz = torch.rand(1, 3, 10)
    with pytest.raises(RuntimeError):
        oracle.farthest_point_sample(xyz, 11)      # M > N          (sampling_kernel.cu:137)
    with pytest.raises(RuntimeError):
        oracle.farthest_point_sample(xyz, 0)       # M <= 0         (sampling_kernel.cu:136)
    with pytest.raises(RuntimeError):
        oracle.point_search(xyz, xyz[:, :, :2], 3)  # Nk < 3        (interpolate_kernel.cu:102)
    with pytest.raises(RuntimeError):
        oracle.point_search(xyz, xyz, 2)           # k != 3         (interpolate_kernel.cu:101)
    idx, cnt = oracle.ball_query(xyz, xyz + 10.0, 0.1, 4)   # nothing in range -> zeros, count 0
    assert idx.abs().sum() == 0 and cnt.sum() == 0
    idx, cnt = oracle.ball_query(xyz, xyz, 1e-3, 4)         # only itself -> replicated
    assert torch.equal(idx, torch.arange(10).view(1, 10, 1).expand(1, 10, 4)) and (cnt == 1).all()


from helpers import region_net_fixture as _region_net_fixture  # noqa: E402


def test_region_net_oracle_vs_reference_python():
    """oracle/region_oracle.region_net_forward against the fixture the REAL GripperRegionNetwork produced on CPU
    (rows R3-R7 of SURVEY.md section 8a, inference call): decode, closing-box membership + index mapping, refine
    selection.  Index outputs bit-exact, grasp parameters to fp32 round-off."""
    import torch
    from oracle import region_oracle
    ref, _, sd, inp = _region_net_fixture()
    out = region_oracle.region_net_forward(sd, inp, [float(x) for x in ref["params"]])
    np.testing.assert_allclose(out["next_grasp"].numpy(), ref["next_grasp"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(out["gripper_mask"].numpy(), ref["gripper_mask"])
    assert np.array_equal(out["gripper_pc_index"].numpy(), ref["gripper_pc_index"])
    assert np.array_equal(out["gripper_pc_index_inall"].numpy(), ref["gripper_pc_index_inall"])
    assert np.array_equal(out["final_mask"].numpy(), ref["final_mask"])
    assert np.array_equal(out["final_mask_sthre"].numpy(), ref["final_mask_sthre"])
    for k in ("sel_class", "sel_score", "sel_stage2"):
        np.testing.assert_allclose(out[k].numpy(), ref[k], rtol=1e-4, atol=1e-5)
    assert 0 < len(ref["final_mask_sthre"]) < len(ref["final_mask"]) < len(ref["gripper_mask"]) < 12   # every branch is hit
