"""GPU parity, operators: every call goes torch tensor -> pn2_ext shim -> C ABI -> sm_100a kernel and is compared
with the CPU oracle on the same seeded inputs (bit-exact for indices / counts / squared distances) and with the
golden vectors recorded from the reference's own CUDA kernels."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden
from helpers import bq_rowhash

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ext(lib_path):
    from regnet_for_3d_grasping_b200 import pn2_ext
    assert torch.cuda.is_available()
    return pn2_ext


def _cases():
    from oracle import gen_golden_gpu
    return gen_golden_gpu.cases()


@pytest.mark.parametrize("name", list(_cases().keys()))
def test_search_ops_bit_exact_vs_oracle(ext, oracle, name):
    c = _cases()[name]
    pc = torch.from_numpy(c["pts"])
    xyz_c = pc[:, :, :3].permute(0, 2, 1)
    xyz = pc.cuda()[:, :, :3].permute(0, 2, 1)         # non-contiguous (B,3,N) view, stride 6, like score_network.py:46
    B = xyz.shape[0]
    idx = ext.farthest_point_sample(xyz, c["M"])
    want = oracle.farthest_point_sample(xyz_c, c["M"])
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (B, c["M"])
    assert torch.equal(idx.cpu(), want), f"{name}: FPS mismatch at {(idx.cpu() != want).nonzero()[:3].tolist()}"
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, c["M"]))
    bq, cnt = ext.ball_query(xyz, new_xyz, c["radius"], c["K"])
    wbq, wcnt = oracle.ball_query(xyz_c, new_xyz.cpu(), c["radius"], c["K"])
    assert bq.dtype == torch.int64 and cnt.dtype == torch.int64
    assert torch.equal(cnt.cpu(), wcnt) and torch.equal(bq.cpu(), wbq), f"{name}: ball query mismatch"
    if c["M"] >= 3:
        nn, nnd = ext.point_search(xyz, new_xyz, 3)
        wnn, wnnd = oracle.point_search(xyz_c, new_xyz.cpu(), 3)
        assert torch.equal(nn.cpu(), wnn), f"{name}: 3-NN index mismatch"
        assert torch.equal(nnd.cpu(), wnnd), f"{name}: 3-NN squared distance mismatch (must be bit-equal)"


def test_search_ops_vs_reference_cuda_golden(ext):
    ref = golden("ref_cuda_ops.npz")
    for name, c in _cases().items():
        xyz = torch.from_numpy(c["pts"]).cuda()[:, :, :3].permute(0, 2, 1)
        B = xyz.shape[0]
        idx = ext.farthest_point_sample(xyz, c["M"])
        assert np.array_equal(idx.cpu().numpy(), ref[name + ".fps"]), name
        new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, c["M"]))
        bq, cnt = ext.ball_query(xyz, new_xyz, c["radius"], c["K"])
        assert np.array_equal(cnt.cpu().numpy(), ref[name + ".bqcnt"]), name
        if name + ".bq" in ref.files:
            assert np.array_equal(bq.cpu().numpy(), ref[name + ".bq"]), name
        else:
            assert np.array_equal(bq_rowhash(bq.cpu().numpy(), c["K"]), ref[name + ".bq_rowhash"]), name
        if c["M"] >= 3:
            nn, nnd = ext.point_search(xyz, new_xyz, 3)
            if name + ".nn" in ref.files:
                assert np.array_equal(nn.cpu().numpy(), ref[name + ".nn"]) and np.array_equal(nnd.cpu().numpy(), ref[name + ".nnd"]), name
            else:
                assert np.array_equal(nn.cpu().numpy()[:, :4096], ref[name + ".nn_head"]), name


@pytest.mark.parametrize("cs,threads", [(1, 128), (4, 128), (8, 128), (1, 256), (2, 256), (4, 256), (8, 256), (1, 512), (2, 512), (4, 512), (8, 512),
                                        (1, 1024), (2, 1024), (4, 1024), (8, 1024), (2, -512), (8, -256)])
def test_fps_every_cluster_configuration(lib_path, oracle, cs, threads):
    """All launch shapes of the cluster kernel give the same (reference) answer, contiguous planar input too."""
    from regnet_for_3d_grasping_b200 import _lib, synth
    lib = _lib.load()
    for kind, n, m in (("lattice", 5120, 700), ("table", 7000, 900), ("cube", 300, 300), ("lattice", 1500, 400),
                       ("lattice", 300, 120)):
        pts = synth.batch(kind, [n, n + 1], n)
        xyz_c = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1).contiguous()
        want = oracle.farthest_point_sample(xyz_c, m)
        x = xyz_c.cuda()
        idx = torch.empty(2, m, dtype=torch.int64, device="cuda")
        idx32 = torch.empty(2, m, dtype=torch.int32, device="cuda")
        nx = torch.empty(2, 3, m, device="cuda")
        rc = lib.regnet_farthest_point_sample_ex(ctypes.c_void_p(x.data_ptr()), x.stride(0), x.stride(1), x.stride(2), 2, n, m,
                                                 ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(idx32.data_ptr()),
                                                 ctypes.c_void_p(nx.data_ptr()), cs, threads, None)
        assert rc == 0, lib.regnet_last_error()
        torch.cuda.synchronize()
        assert torch.equal(idx.cpu(), want), (kind, cs, threads)
        assert torch.equal(idx32.cpu().long(), want)
        assert torch.equal(nx.cpu(), xyz_c.gather(2, want.unsqueeze(1).expand(2, 3, m)))


def test_fps_full_size_cloud_matches_oracle(ext, oracle):
    """BASELINE config size: 25 600 points -> 5 120 centroids (one cloud; the oracle needs ~2 s for it)."""
    from regnet_for_3d_grasping_b200 import synth
    pts = synth.batch("table", [77], 25600)
    want = oracle.farthest_point_sample(torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1), 5120)
    got = ext.farthest_point_sample(torch.from_numpy(pts).cuda()[:, :, :3].permute(0, 2, 1), 5120)
    assert torch.equal(got.cpu(), want)
    assert got.unique().numel() > 5000   # FPS picks (nearly) distinct points; duplicates only where the cloud has them


def test_group_and_interpolate_forward_backward(ext, oracle):
    g = torch.Generator().manual_seed(5)
    B, C, N, M, K = 2, 19, 1024, 256, 16
    from regnet_for_3d_grasping_b200 import synth
    pts = synth.batch("cube", [8, 9], N)
    xyz_c = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1)
    idx = oracle.farthest_point_sample(xyz_c, M)
    new_xyz = xyz_c.gather(2, idx.unsqueeze(1).expand(B, 3, M))
    bq, _ = oracle.ball_query(xyz_c, new_xyz, 0.12, K)
    nn, nnd = oracle.point_search(xyz_c, new_xyz, 3)
    inv = 1.0 / torch.clamp(nnd, min=1e-10)
    w = inv / inv.sum(2, keepdim=True)
    feat = torch.randn(B, C, N, generator=g)
    feat_nc = torch.randn(B, N, C, generator=g).permute(0, 2, 1)      # strided input
    for f in (feat, feat_nc):
        got = ext.group_points_forward(f.cuda(), bq.cuda())
        assert torch.equal(got.cpu(), oracle.group_points_forward(f, bq))
    gout = torch.randn(B, C, M, K, generator=g)
    np.testing.assert_allclose(ext.group_points_backward(gout.cuda(), bq.cuda(), N).cpu().numpy(),
                               oracle.group_points_backward(gout, bq, N).numpy(), rtol=1e-5, atol=1e-5)
    sfeat = torch.randn(B, C, M, generator=g)
    got = ext.interpolate_forward(sfeat.cuda(), nn.cuda(), w.cuda())
    assert torch.equal(got.cpu(), oracle.interpolate_forward(sfeat, nn, w)), "same fma chain => bit equal"
    iout = torch.randn(B, C, N, generator=g)
    np.testing.assert_allclose(ext.interpolate_backward(iout.cuda(), nn.cuda(), w.cuda(), M).cpu().numpy(),
                               oracle.interpolate_backward(iout, nn, w, M).numpy(), rtol=1e-5, atol=1e-5)
    ext.check_index_errors()
    bad = bq.clone()
    bad[0, 0, 0] = N + 5
    ext.group_points_forward(feat.cuda(), bad.cuda())
    with pytest.raises(RuntimeError, match="out of range"):
        ext.check_index_errors()


def test_autograd_functions_on_gpu(ext, oracle):
    from regnet_for_3d_grasping_b200 import function as F
    g = torch.Generator().manual_seed(9)
    xyz = torch.rand(2, 3, 256, generator=g).cuda()
    feat = torch.randn(2, 6, 256, generator=g).cuda().requires_grad_(True)
    idx = F.farthest_point_sample(xyz, 64)
    new_xyz = F.gather_points(xyz, idx)
    nbr, _ = F.ball_query(xyz, new_xyz, 0.3, 8)
    grouped = F.group_points(feat, nbr)
    grouped.square().sum().backward()
    ref = torch.zeros_like(feat)
    ref.scatter_add_(2, nbr.view(2, 1, -1).expand(2, 6, -1), (2 * grouped.detach()).view(2, 6, -1))
    np.testing.assert_allclose(feat.grad.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_bad_arguments_raise_like_the_reference(ext):
    """The reference's TORCH_CHECK / CHECK_* failures (sampling_kernel.cu:130-137, interpolate_kernel.cu:97-105) surface as
    RuntimeError here too; shapes that are legal there are legal here."""
    x = torch.rand(1, 3, 10, device="cuda")
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(x, 11)
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(x, 0)
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(torch.rand(1, 4, 10, device="cuda"), 2)
    with pytest.raises(RuntimeError):
        ext.point_search(x, x, 2)
    with pytest.raises(RuntimeError):
        ext.point_search(x, x[:, :, :2], 3)
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(x.cpu(), 2)              # CHECK_CUDA
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(x.half(), 2)             # only float / double are dispatched
    idx, cnt = ext.ball_query(x, x + 10.0, 0.1, 4)
    assert idx.abs().sum().item() == 0 and cnt.sum().item() == 0
    assert ext.farthest_point_sample(torch.rand(0, 3, 10, device="cuda"), 2).shape == (0, 2)


def test_sizes_beyond_the_tuned_kernels_fall_back(ext, oracle):
    """The reference has no size limits: more than 128 neighbours per ball and clouds of more than 65 536 points take
    generic kernels here and still return the oracle's bits."""
    from regnet_for_3d_grasping_b200 import synth
    pts = synth.batch("table", [41], 6000)
    xyz_c = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1)
    xyz = torch.from_numpy(pts).cuda()[:, :, :3].permute(0, 2, 1)
    ctr_c = xyz_c[:, :, ::40].contiguous()
    bq, cnt = ext.ball_query(xyz, ctr_c.cuda(), 0.25, 200)                      # K = 200 > 128
    wbq, wcnt = oracle.ball_query(xyz_c, ctr_c, 0.25, 200)
    assert torch.equal(bq.cpu(), wbq) and torch.equal(cnt.cpu(), wcnt)
    assert int(cnt.max()) > 128
    big = synth.batch("table", [42], 70000)                                    # N = 70 000 > 65 536
    big_c = torch.from_numpy(big[:, :, :3]).permute(0, 2, 1)
    got = ext.farthest_point_sample(torch.from_numpy(big).cuda()[:, :, :3].permute(0, 2, 1), 96)
    assert torch.equal(got.cpu(), oracle.farthest_point_sample(big_c, 96))


def _ref_double_ext():
    """The reference's own extension (oracle/_ref, built in the build container, travels with the snapshot): its double
    kernels are the ground truth for the float64 path.  None when it is not on this box."""
    try:
        from oracle import build_ref
        return build_ref.load()
    except Exception:
        return None


def test_float64_operators(ext):
    """float64 inputs (the reference instantiates float and double): search operators against the reference's own double
    kernels when oracle/_ref is present, else against a float64 torch restatement; value operators against torch."""
    from regnet_for_3d_grasping_b200 import synth
    g = torch.Generator().manual_seed(77)
    pts = torch.from_numpy(synth.batch("cube", [51, 52], 700)[:, :, :3]).double()
    pts = pts + torch.rand(pts.shape, generator=g, dtype=torch.float64) * 1e-9       # not representable in float32
    xyz = pts.cuda().permute(0, 2, 1)                                               # strided (B,3,N)
    M, K, r = 200, 24, 0.22
    idx = ext.farthest_point_sample(xyz, M)
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(2, 3, M))
    bq, cnt = ext.ball_query(xyz, new_xyz, r, K)
    nn, nnd = ext.point_search(xyz, new_xyz, 3)
    assert nnd.dtype == torch.float64 and idx.dtype == torch.int64
    ref = _ref_double_ext()
    if ref is not None:
        xc = xyz.contiguous()
        assert torch.equal(idx, ref.farthest_point_sample(xc, M))
        rbq, rcnt = ref.ball_query(xc, new_xyz.contiguous(), r, K)
        assert torch.equal(bq, rbq) and torch.equal(cnt, rcnt)
        rnn, rnnd = ref.point_search(xc, new_xyz.contiguous(), 3)
        assert torch.equal(nn, rnn) and torch.equal(nnd, rnnd)
    # float64 restatement in torch (greedy FPS, brute-force ball query / 3-NN)
    P = pts.cuda()
    d2 = ((P.unsqueeze(2) - P.unsqueeze(1)) ** 2)
    D = d2[..., 1] + d2[..., 0] + d2[..., 2]
    for b in range(2):
        md = torch.full((700,), float("inf"), dtype=torch.float64, device="cuda")
        cur, picks = 0, [0]
        for _ in range(M - 1):
            md = torch.minimum(md, D[b, cur])
            cur = int(torch.argmax(md))
            picks.append(cur)
        same = (torch.tensor(picks, device="cuda") == idx[b]).float().mean().item()
        assert same > 0.99, same          # identical up to exact ties (the reference's tie order is not argmax's)
    Dq = ((new_xyz.permute(0, 2, 1).unsqueeze(2) - P.unsqueeze(1)) ** 2)
    Dq = Dq[..., 1] + Dq[..., 0] + Dq[..., 2]
    assert torch.equal(cnt, (Dq < r * r).sum(-1).clamp(max=K))
    Dn = ((P.unsqueeze(2) - new_xyz.permute(0, 2, 1).unsqueeze(1)) ** 2)          # queries = the cloud, keys = the centroids
    Dn = Dn[..., 1] + Dn[..., 0] + Dn[..., 2]
    want_nn = Dn.topk(3, dim=-1, largest=False)
    assert torch.allclose(nnd, want_nn.values, rtol=1e-12, atol=1e-18)
    assert (nn == want_nn.indices).float().mean().item() > 0.999
    # value operators: exact gathers, double scatter-adds
    feat = torch.randn(2, 5, 700, generator=g, dtype=torch.float64).cuda()
    grouped = ext.group_points_forward(feat, bq)
    assert grouped.dtype == torch.float64 and torch.equal(grouped[1, 3, 7, 2], feat[1, 3, bq[1, 7, 2]])
    gin = ext.group_points_backward(torch.ones_like(grouped), bq, 700)
    assert torch.allclose(gin.sum(), torch.tensor(float(grouped.numel()), dtype=torch.float64, device="cuda"))
    sfeat = torch.randn(2, 5, M, generator=g, dtype=torch.float64).cuda()       # features at the M centroids (the keys)
    w = torch.rand(2, 700, 3, generator=g, dtype=torch.float64).cuda()
    out = ext.interpolate_forward(sfeat, nn, w)
    want = sum(sfeat.gather(2, nn[:, :, k].unsqueeze(1).expand(2, 5, 700)) * w[:, :, k].unsqueeze(1) for k in range(3))
    assert out.dtype == torch.float64 and torch.allclose(out, want, rtol=1e-14)
    back = ext.interpolate_backward(torch.ones(2, 5, 700, dtype=torch.float64, device="cuda"), nn, w, M)
    assert tuple(back.shape) == (2, 5, M)
    assert torch.allclose(back.sum(dim=2), w.sum(dim=(1, 2)).unsqueeze(1).expand(2, 5), rtol=1e-12)


def test_dgcnn_alias(ext):
    from regnet_for_3d_grasping_b200 import dgcnn_ext
    torch.manual_seed(1)
    feat = torch.rand(2, 4, 5, device="cuda")
    knn = torch.randint(0, 5, (2, 5, 3), device="cuda")
    want = torch.gather(feat.unsqueeze(2).expand(2, 4, 5, 5), 3, knn.unsqueeze(1).expand(2, 4, 5, 3))
    assert torch.equal(dgcnn_ext.gather_knn_forward(feat, knn), want)        # functions/gather_knn.py:26-55 self-check
    g = dgcnn_ext.gather_knn_backward(torch.ones(2, 4, 5, 3, device="cuda"), knn)
    ref = torch.zeros(2, 4, 5, device="cuda").scatter_add_(2, knn.view(2, 1, 15).expand(2, 4, 15), torch.ones(2, 4, 15, device="cuda"))
    assert torch.allclose(g, ref)


@pytest.mark.parametrize("kind,N,M,radius", [("table", 25600, 5120, 0.02), ("table", 8192, 2048, 0.4),
                                            ("lattice", 5120, 1024, 0.08), ("cube", 4096, 1024, 0.1)])
def test_grid_and_brute_force_operator_paths_agree(ext, monkeypatch, kind, N, M, radius):
    """pn2_ext.ball_query / point_search route big clouds through the uniform-grid kernels (regnet_ball_query_ws /
    regnet_point_search_ws); REGNET_API_BRUTE forces the brute-force scan.  Both must give the same bits, including
    the dense-ball overflow path (radius 0.4 on a table cloud: thousands of hits per ball) and keys == queries."""
    from regnet_for_3d_grasping_b200 import synth
    xyz = torch.from_numpy(synth.batch(kind, [21, 22], N)).cuda()[:, :, :3].permute(0, 2, 1)
    idx = ext.farthest_point_sample(xyz, M)
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(2, 3, M))
    got = ext.ball_query(xyz, new_xyz, radius, 64) + ext.point_search(new_xyz, xyz, 3) + ext.point_search(xyz, xyz, 3)
    monkeypatch.setenv("REGNET_API_BRUTE", "1")
    want = ext.ball_query(xyz, new_xyz, radius, 64) + ext.point_search(new_xyz, xyz, 3) + ext.point_search(xyz, xyz, 3)
    for g, w, what in zip(got, want, ["bq index", "bq count", "nn index", "nn dist", "self nn index", "self nn dist"]):
        assert torch.equal(g, w), f"{kind} N={N}: {what} differs between the grid and the brute-force path"
    assert int(got[1].min()) >= 1   # every centroid is one of the points


@pytest.mark.parametrize("force_global", [False, True])
def test_scatter_add_backward_both_paths(ext, oracle, monkeypatch, force_global):
    """group_points_backward / interpolate_backward: the shared-memory row-accumulating kernels (default when the row
    fits) and the global-atomics kernels (REGNET_SCATTER_GLOBAL, or big rows) against the oracle, at an SA-level-1-like
    shape with under-full balls (duplicated neighbours => many contributions per source point)."""
    if force_global:
        monkeypatch.setenv("REGNET_SCATTER_GLOBAL", "1")
    from regnet_for_3d_grasping_b200 import synth
    g = torch.Generator().manual_seed(15)
    B, C, N, M, K = 2, 24, 5120, 1024, 64
    pts = synth.batch("table", [31, 32], N)
    xyz_c = torch.from_numpy(pts[:, :, :3]).permute(0, 2, 1)
    idx = oracle.farthest_point_sample(xyz_c, M)
    new_xyz = xyz_c.gather(2, idx.unsqueeze(1).expand(B, 3, M))
    bq, _ = oracle.ball_query(xyz_c, new_xyz, 0.03, K)
    gout = torch.randn(B, C, M, K, generator=g)
    got = ext.group_points_backward(gout.cuda(), bq.cuda(), N).cpu()
    want = oracle.group_points_backward(gout, bq, N)
    assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    nn, nnd = oracle.point_search(xyz_c, new_xyz, 3)
    inv = 1.0 / torch.clamp(nnd, min=1e-10)
    w = inv / inv.sum(2, keepdim=True)
    iout = torch.randn(B, C, N, generator=g)
    got = ext.interpolate_backward(iout.cuda(), nn.cuda(), w.cuda(), M).cpu()
    want = oracle.interpolate_backward(iout, nn, w, M)
    assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    ext.check_index_errors()
