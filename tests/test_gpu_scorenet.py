"""GPU parity, the fused ScoreNet forward (native plan) against the oracle restatement of the reference
(oracle/ref_modules.py, CPU fp64 features / fp32 search ops) and the committed golden fixture."""
import numpy as np
import pytest
import torch

from conftest import golden
from helpers import assert_features_close

pytestmark = pytest.mark.gpu

SMALL = dict(num_centroids=(512, 128, 32), radius=(0.05, 0.12, 0.4))


def _oracle_forward(oracle, sd, pc, **arch):
    from oracle import ref_modules
    with torch.no_grad():
        return ref_modules.scorenet_forward(sd, pc, oracle.as_pn2_ext(), dtype=torch.float64, keep=True, **arch)


@pytest.mark.parametrize("engine", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("kind", ["table", "lattice"])
def test_plan_small_cloud_every_stage(lib_path, oracle, engine, kind):
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    B, N = 2, 2048
    pc = torch.from_numpy(synth.batch(kind, [31, 32], N))
    if kind == "lattice":
        pc[:, :, :3] *= 4.0
    sd = ref_modules.random_scorenet_state(seed=5)
    feat, score, dbg = _oracle_forward(oracle, sd, pc, **SMALL)
    plan = ScoreNetPlan(B, N, "cuda", engine=engine, **SMALL)
    plan.bind_state(sd)
    got_feat, got_score = plan.forward(pc.cuda())
    torch.cuda.synchronize()
    M = SMALL["num_centroids"]
    nd = (M[1], M[0], N)
    for i in range(3):
        assert torch.equal(plan.intermediate(f"fps{i}", torch.int32, (B, M[i])).cpu().long(), dbg[f"fps{i}"]), f"fps{i}"
        assert torch.equal(plan.intermediate(f"bq{i}", torch.int32, (B, M[i], 64)).cpu().long(), dbg[f"bq{i}"]), f"bq{i}"
        assert torch.equal(plan.intermediate(f"nn{i}", torch.int32, (B, nd[i], 3)).cpu().long(), dbg[f"nn{i}"]), f"nn{i}"
    for i, c in enumerate((256, 512, 1024)):
        assert_features_close(plan.intermediate(f"sa{i}", torch.float32, (B, M[i], c)), dbg[f"sa{i}"].transpose(1, 2), what=f"sa{i}")
    for i, c in enumerate((1024, 512)):
        assert_features_close(plan.intermediate(f"fp{i}", torch.float32, (B, nd[i], c)), dbg[f"fp{i}"].transpose(1, 2), what=f"fp{i}")
    assert_features_close(got_feat, feat, what="all_feature")
    assert_features_close(got_score, score, what="score")
    assert plan.launch_count > 30
    # same input again -> identical bits (no atomics / races in the forward)
    f2, s2 = plan.forward(pc.cuda())
    torch.cuda.synchronize()
    assert torch.equal(f2, got_feat) and torch.equal(s2, got_score)
    plan.close()


def test_plan_vs_reference_python_golden(lib_path):
    """N=6144 with the reference's real architecture constants, against outputs of the reference's own modules."""
    ref = golden("ref_py_scorenet_n6144.npz")
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    pc = torch.from_numpy(synth.batch("table", [11], 6144)).cuda()
    plan = ScoreNetPlan(1, 6144, "cuda")
    plan.bind_state(ref_modules.random_scorenet_state(seed=3))
    feat, score = plan.forward(pc)
    torch.cuda.synchronize()
    for i, m in enumerate((5120, 1024, 256)):
        assert np.array_equal(plan.intermediate(f"fps{i}", torch.int32, (1, m)).cpu().numpy()[0], ref[f"fps{i}"])
        assert np.array_equal(plan.intermediate(f"bq{i}", torch.int32, (1, m, 64)).cpu().numpy()[0].sum(1), ref[f"bq{i}_sum"])
    assert np.array_equal(plan.intermediate("nn2", torch.int32, (1, 6144, 3)).cpu().numpy()[0, ::13], ref["nn2"])
    assert_features_close(feat[0, torch.from_numpy(ref["rows"]).long().cuda()], ref["all_feature_rows"], what="all_feature rows")
    assert_features_close(score[0], ref["score"], what="score")


def test_dropin_scorenetwork_fused_equals_module_path(lib_path):
    """multi_model.score_network.ScoreNetwork drop-in: eval-mode fused plan == its own op-by-op module path
    (torch conv/BN over this package's point operators) == golden; loss path as in the reference."""
    ref = golden("ref_py_scorenet_n6144.npz")
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = ScoreNetwork(training=True).cuda().eval()
    net.load_state_dict(ref_modules.random_scorenet_state(seed=3))
    pc = torch.from_numpy(synth.batch("table", [11], 6144)).cuda()
    tgt = torch.from_numpy(synth.scores_like_dataset(5, 1, 6144)).cuda()
    with torch.no_grad():
        feat, score, loss = net(pc, tgt)
        body = net.extrat_featurePN2
        f2, s2 = body._forward_modules(pc[:, :, :6].permute(0, 2, 1))
    assert feat.shape == (1, 6144, 256) and score.shape == (1, 6144)
    assert_features_close(feat, f2.transpose(2, 1), what="fused vs module path: all_feature")
    assert_features_close(score, s2, what="fused vs module path: score")
    assert_features_close(score[0], ref["score"], what="score vs golden")
    np.testing.assert_allclose(loss.item(), golden("ref_py_scorenet_loss.npz")["loss"], rtol=1e-4)
    # parameters change -> the plan re-folds them
    with torch.no_grad():
        net.extrat_featurePN2.bn_score.bias.add_(0.5)
        _, s3, _ = net(pc)
    assert (s3 - score).abs().max().item() > 1e-3


def test_train_mode_step_runs_and_updates(lib_path):
    """train.py --mode pretrain_score shape of a step (train.py:143-149): forward, MSE loss, backward, Adam."""
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
    torch.manual_seed(0)
    net = ScoreNetwork(training=True).cuda().train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    pc = torch.from_numpy(synth.batch("table", [1], 5632)).cuda()
    tgt = torch.from_numpy(synth.scores_like_dataset(2, 1, 5632)).cuda()
    before = net.extrat_featurePN2.sa_modules[0].mlp[0].conv.weight.detach().clone()
    _, score, loss = net(pc, tgt)
    loss.sum().backward()
    opt.step()
    assert torch.isfinite(loss).all() and score.shape == (1, 5632)
    assert (net.extrat_featurePN2.sa_modules[0].mlp[0].conv.weight.detach() - before).abs().max().item() > 0


def test_full_size_batch_properties(lib_path, oracle):
    """BASELINE config 2 (B=15 x 25 600): size-independent properties + exact FPS parity on one of the clouds."""
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    B, N = 15, 25600
    pts = synth.batch("table", range(100, 100 + B), N)
    pc = torch.from_numpy(pts).cuda()
    plan = ScoreNetPlan(B, N, "cuda")
    plan.bind_state(ref_modules.random_scorenet_state(seed=0))
    feat, score = plan.forward(pc)
    torch.cuda.synchronize()
    assert torch.isfinite(feat).all() and torch.isfinite(score).all()
    assert (score > 0).all() and (score < 1).all() and (feat >= 0).all()       # sigmoid / ReLU ranges
    fps0 = plan.intermediate("fps0", torch.int32, (B, 5120)).cpu()
    assert (fps0[:, 0] == 0).all() and fps0.min() >= 0 and fps0.max() < N
    want = oracle.farthest_point_sample(torch.from_numpy(pts[7:8, :, :3]).permute(0, 2, 1), 5120)
    assert torch.equal(fps0[7:8].long(), want)
    bq0 = plan.intermediate("bq0", torch.int32, (B, 5120, 64)).cpu()
    assert (bq0[:, :, 0] >= 0).all() and (bq0[:, :, 1:] >= bq0[:, :, :1]).all()   # first hit is the smallest index
    # clouds are independent units: cloud 3 alone gives the same rows (batch-axis sharding is exact)
    plan1 = ScoreNetPlan(1, N, "cuda")
    plan1.bind_state(ref_modules.random_scorenet_state(seed=0))
    f1, s1 = plan1.forward(pc[3:4].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(f1[0], feat[3]) and torch.equal(s1[0], score[3])


def test_full_size_batch_features_match_the_reference_gpu_path(lib_path):
    """BASELINE configs[1] itself (B = 15 x 25 600): all_feature and the scores of the fused plan against the forward the
    way the reference computes it on a GPU -- the reference's OWN CUDA kernels (oracle/_ref/pn2_ext_ref.so, built from the
    reference's .cu files; travels with the repository snapshot) under the restated modules with torch's convolutions in
    strict fp32.  Without the reference extension on the box, this repo's operators (bit-identical to it, see
    test_gpu_ops.py) stand in under the same fp32 torch modules.  Indices exact, features within 1e-4 (helpers.py)."""
    from helpers import assert_features_close
    from oracle import ref_modules
    from regnet_for_3d_grasping_b200 import pn2_ext, synth, weights
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    try:
        from oracle import build_ref
        ext = build_ref.load()
        which = "reference CUDA kernels (oracle/_ref)"
    except Exception:
        ext, which = pn2_ext, "this repo's operators (oracle/_ref not on this box)"
    B, N = 15, 25600
    pc = torch.from_numpy(synth.batch("table", range(100, 100 + B), N)).cuda()
    sd = {k: v.cuda() for k, v in weights.random_scorenet_state(seed=0).items()}
    plan = ScoreNetPlan(B, N, "cuda")
    plan.bind_state(sd)
    feat, score = plan.forward(pc)
    torch.cuda.synchronize()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            rf, rs, dbg = ref_modules.scorenet_forward(sd, pc, ext, keep=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    for lvl, m in enumerate((5120, 1024, 256)):
        assert torch.equal(plan.intermediate(f"fps{lvl}", torch.int32, (B, m)).long(), dbg[f"fps{lvl}"].long()), (which, lvl)
        assert torch.equal(plan.intermediate(f"bq{lvl}", torch.int32, (B, m, 64)).long(), dbg[f"bq{lvl}"].long()), (which, lvl)
    err = ((feat - rf).abs().max() / rf.abs().max()).item()
    print(f"full-size all_feature vs {which}: max rel err {err:.3e}; score max abs err {(score - rs).abs().max().item():.3e}")
    rows = torch.arange(0, N, 37, device="cuda")
    for b in range(B):
        assert_features_close(feat[b][rows], rf[b][rows], what=f"all_feature, cloud {b} ({which})")
    assert err < 1e-4 and (score - rs).abs().max().item() < 1e-4


def test_prefetch_pipeline_equals_plain_forward(lib_path):
    """Software-pipelined throughput mode (prefetch geometry of batch i+1 during the MLPs of batch i) returns
    exactly what independent forwards return, for alternating different inputs and both slot orders."""
    from regnet_for_3d_grasping_b200 import synth, weights
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    B, N = 2, 6144
    sd = weights.random_scorenet_state(seed=4)
    batches = [torch.from_numpy(synth.batch("table", [50 + 2 * k, 51 + 2 * k], N)).cuda() for k in range(3)]
    plain = ScoreNetPlan(B, N, "cuda")
    plain.bind_state(sd)
    want = []
    for pc in batches:
        f, s = plain.forward(pc)
        torch.cuda.synchronize()
        want.append((f.clone(), s.clone()))
    plan = ScoreNetPlan(B, N, "cuda")
    plan.bind_state(sd)
    order = [0, 1, 2, 1, 0, 0, 2]
    plan.prefetch(batches[order[0]])
    got = []
    for i, k in enumerate(order):
        if i + 1 < len(order):
            plan.prefetch(batches[order[i + 1]])
        f, s = plan.forward(batches[k])
        got.append((k, f, s))          # fresh output tensors per call
    torch.cuda.synchronize()
    for k, f, s in got:
        assert torch.equal(f, want[k][0]) and torch.equal(s, want[k][1]), f"pipelined result differs for batch {k}"
    # a forward without matching prefetch still works (computes geometry inline) ...
    f, s = plan.forward(batches[1])
    torch.cuda.synchronize()
    assert torch.equal(s, want[1][1])
    # ... and a third outstanding prefetch is refused
    plan.prefetch(batches[0])
    plan.prefetch(batches[1])
    with pytest.raises(RuntimeError, match="prefetch"):
        plan.prefetch(batches[2])
    f, s = plan.forward(batches[0])
    f2, s2 = plan.forward(batches[1])
    torch.cuda.synchronize()
    assert torch.equal(s, want[0][1]) and torch.equal(s2, want[1][1])


@pytest.mark.parametrize("B,N", [(15, 25600), (11, 25000)])
def test_warp_producers_and_stream_modes_give_identical_bits(lib_path, monkeypatch, B, N):
    """The warp-cooperative gather-affine / interpolate-affine producers compute the same fma chains as the
    thread-per-(row, 4 channels) kernels they replace (REGNET_AFFINE_V1=1 selects those), and the three-stream
    orchestration (side mode 3) only reorders launches: all variants must agree bit for bit -- at the BASELINE batch and
    at a ragged one (11 x 25 000: B*N is not a multiple of 32, so the interpolate producer's tail path runs)."""
    from regnet_for_3d_grasping_b200 import synth, weights
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    sd = weights.random_scorenet_state(seed=6)
    pc = torch.from_numpy(synth.batch("table", range(300, 300 + B), N)).cuda()
    monkeypatch.setenv("REGNET_SA_FUSED_A", "0")     # keep the set-abstraction producers in the comparison

    def run(side_stream):
        plan = ScoreNetPlan(B, N, "cuda", side_stream=side_stream)
        plan.bind_state(sd)
        f, s = plan.forward(pc)
        torch.cuda.synchronize()
        plan.close()
        return f, s

    f3, s3 = run(3)
    f0, s0 = run(0)
    assert torch.equal(f3, f0) and torch.equal(s3, s0), "side-stream mode 3 differs from the single-stream forward"
    monkeypatch.setenv("REGNET_AFFINE_V1", "1")
    f1, s1 = run(3)
    assert torch.equal(f3, f1) and torch.equal(s3, s1), "warp-cooperative producers differ from the per-thread producers"


@pytest.mark.parametrize("B,N", [(15, 25600), (3, 6144)])
def test_fused_operand_layers_match_the_materialised_ones(lib_path, B, N):
    """Set-abstraction levels 1 and 2, second layer: the operand relu(Z'[g] + T) built inside the GEMM (gemm_fused_a.cu,
    option sa_fused_a = 1) against the same operand materialised by a gather-add pass and fed to the plain GEMM (= 3): the
    two feed identical bits to identical MMA sequences, so every output must be EQUAL.  Against the older form that keeps
    the xyz term as W_x (xyz - centre) in the producer (= 0) the folded form W_x xyz - W_x centre differs by a few ulps of
    the operand: 2e-5 on the features."""
    from regnet_for_3d_grasping_b200 import synth, weights
    from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
    sd = weights.random_scorenet_state(seed=8)
    pc = torch.from_numpy(synth.batch("table", range(400, 400 + B), N)).cuda()
    out = {}
    for mode in (1, 3, 0):
        plan = ScoreNetPlan(B, N, "cuda")
        plan.set_option("sa_fused_a", mode)
        plan.bind_state(sd)
        f, s = plan.forward(pc)
        torch.cuda.synchronize()
        names = [label for label, _ in plan.profile_forward(pc)]
        assert ("sa_fold.1" in names) == (mode != 0) and ("sa_operand.1" in names) == (mode != 1), (mode, names)
        out[mode] = (f.clone(), s.clone())
        plan.close()
    assert torch.equal(out[1][0], out[3][0]) and torch.equal(out[1][1], out[3][1]), "fused and materialised operands differ"
    scale = float(out[0][0].abs().max())
    assert float((out[1][0] - out[0][0]).abs().max()) <= 2e-5 * scale
    assert float((out[1][1] - out[0][1]).abs().max()) <= 1e-5
    # default policy: fused for a lone forward, materialised next to a prefetch -- the same bits either way
    plan = ScoreNetPlan(B, N, "cuda")
    plan.bind_state(sd)
    f_alone, _ = plan.forward(pc)
    pc2 = pc.clone()
    plan.prefetch(pc2)
    f_co, _ = plan.forward(pc)
    f_next, _ = plan.forward(pc2)
    torch.cuda.synchronize()
    assert torch.equal(f_alone, out[1][0]) and torch.equal(f_co, f_alone) and torch.equal(f_next, f_alone)
    plan.close()
