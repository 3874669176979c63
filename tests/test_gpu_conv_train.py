"""GPU parity, training-path 1x1 convolutions on the tcgen05 engine (csrc/conv_train.cu) through the C ABI, and the chained
shared MLP built on them (conv_train.py), against torch in float64 -- the floating-point oracle for this path: the
reference's nn/modules/conv.py:24-36,64-76 and mlp.py:95-106 are plain torch modules.
Tolerance: 1e-4 relative (BASELINE.json north_star) for the split-bf16 arithmetic; 2e-2 for the one-pass bf16 mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return ((got.double().cpu() - want.cpu()).abs().max() / want.abs().max().clamp_min(1e-30)).item()


FPROP_SHAPES = [(1, 64, 256, 128), (2, 6, 320, 128), (3, 259, 1280, 256), (1, 128, 1024, 1), (2, 1536, 256, 1024),
                (2, 515, 2056, 256), (4, 128, 8192, 128)]


@pytest.mark.parametrize("B,K,L,rows", FPROP_SHAPES)
def test_conv1x1_forward_and_moments(lib_path, B, K, L, rows):
    from regnet_for_3d_grasping_b200 import conv_train as ct
    g = torch.Generator().manual_seed(B * 1000 + K + L + rows)
    x = torch.randn(B, K, L, generator=g) + 0.3
    w = torch.randn(rows, K, generator=g) / K ** 0.5
    want = torch.einsum("rk,bkl->brl", w.double(), x.double())
    hi, lo = ct.split_planes(x.cuda())
    a_hi, a_lo = ct.split_weight(w.cuda())
    out, mom = ct.conv1x1(hi, lo, a_hi, a_lo, rows, K, want_moments=True, passes=3)
    assert _rel(out, want) < 1e-4
    assert _rel(mom[:, 0], want.sum(dim=(0, 2))) < 1e-4 or want.sum(dim=(0, 2)).abs().max() < 1e-3
    assert _rel(mom[:, 1], (want * want).sum(dim=(0, 2))) < 1e-4
    out1 = ct.conv1x1(hi, lo, a_hi, a_lo, rows, K, passes=1)
    assert _rel(out1, want) < 2e-2
    # dgrad is the same kernel with the transposed weight
    t_hi, t_lo = ct.split_weight(w.cuda(), transpose=True)
    gz = torch.randn(B, rows, L, generator=g)
    g_hi, g_lo = ct.split_planes(gz.cuda())
    dx = ct.conv1x1(g_hi, g_lo, t_hi, t_lo, K, rows, passes=3)
    assert _rel(dx, torch.einsum("rk,brl->bkl", w.double(), gz.double())) < 1e-4


@pytest.mark.parametrize("B,Co,Ci,L", [(1, 128, 128, 256), (2, 128, 6, 320), (3, 256, 259, 1280), (2, 1024, 1536, 256),
                                       (2, 1, 128, 1024), (5, 512, 256, 4104)])
def test_wgrad(lib_path, B, Co, Ci, L):
    from regnet_for_3d_grasping_b200 import conv_train as ct
    g = torch.Generator().manual_seed(B + Co + Ci + L)
    gz = torch.randn(B, Co, L, generator=g)
    x = torch.randn(B, Ci, L, generator=g) + 0.2
    want = torch.einsum("bol,bil->oi", gz.double(), x.double())
    g_hi, g_lo = ct.split_planes(gz.cuda())
    x_hi, x_lo = ct.split_planes(x.cuda())
    dw = ct.wgrad(g_hi, g_lo, x_hi, x_lo, passes=3)
    assert _rel(dw, want) < 1e-4
    assert torch.equal(dw, ct.wgrad(g_hi, g_lo, x_hi, x_lo, passes=3)), "wgrad is not deterministic"
    assert _rel(ct.wgrad(g_hi, g_lo, x_hi, x_lo, passes=1), want) < 2e-2


# every 1x1 convolution of ScoreNet (utils/pointnet2.py:43-46,82): (cin, cout), positions scaled down
SCORENET_LAYERS = [(6, 128), (128, 128), (128, 256), (259, 256), (256, 256), (256, 512), (515, 512), (512, 512), (512, 1024),
                   (1536, 1024), (1024, 1024), (1280, 512), (512, 512), (515, 256), (256, 256), (256, 512), (512, 256),
                   (256, 128), (128, 1)]


@pytest.mark.parametrize("cin,cout", SCORENET_LAYERS)
def test_every_scorenet_layer_forward_dgrad_wgrad(lib_path, cin, cout):
    """Output, input gradient and weight gradient of every SA / FP / seg / score convolution shape against float64,
    with the feature tolerance of tests/helpers.py (1e-4 of the peak, and 1e-4 relative + 1e-4 rms element-wise)."""
    from helpers import assert_features_close
    from regnet_for_3d_grasping_b200 import conv_train as ct
    B, L = 2, 1536
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(B, cin, L, generator=g).clamp_min(-0.5)          # post-ReLU-like: mostly positive, some zeros
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    gz = torch.randn(B, cout, L, generator=g) * 1e-3
    x_hi, x_lo = ct.split_planes(x.cuda())
    g_hi, g_lo = ct.split_planes(gz.cuda())
    w_hi, w_lo = ct.split_weight(w.cuda())
    t_hi, t_lo = ct.split_weight(w.cuda(), transpose=True)
    assert_features_close(ct.conv1x1(x_hi, x_lo, w_hi, w_lo, cout, cin), torch.einsum("oi,bil->bol", w.double(), x.double()),
                          what="fprop")
    assert_features_close(ct.conv1x1(g_hi, g_lo, t_hi, t_lo, cin, cout), torch.einsum("oi,bol->bil", w.double(), gz.double()),
                          what="dgrad")
    assert_features_close(ct.wgrad(g_hi, g_lo, x_hi, x_lo), torch.einsum("bol,bil->oi", gz.double(), x.double()), what="wgrad")


def _mlp_pair(cin, widths, ndim, dropout=0.0, seed=0, relu=True):
    from regnet_for_3d_grasping_b200.nn_layers import SharedMLP
    torch.manual_seed(seed)
    ours = SharedMLP(cin, widths, ndim=ndim, dropout_prob=dropout)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for blk in ours:
            blk.bn.weight.copy_(torch.randn(blk.bn.weight.shape, generator=g))       # both signs
            blk.bn.bias.copy_(torch.randn(blk.bn.bias.shape, generator=g) * 0.3)
            if not relu:
                blk.relu = None
    import copy
    ref = copy.deepcopy(ours).double()
    return ours.cuda().train(), ref.train()


CHAIN_CASES = [((2, 6, 40, 64), (128, 128, 256), True), ((3, 259, 16, 64), (256, 512), True),
               ((2, 515, 1280), (256, 256, 256), False), ((2, 1536, 256), (1024, 1024), False)]


@pytest.mark.parametrize("relu", [False, True])
@pytest.mark.parametrize("shape,widths,pooled", CHAIN_CASES)
def test_mlp_chain_matches_float64_torch(lib_path, monkeypatch, shape, widths, pooled, relu):
    """Values, input gradient and every weight / gamma / beta gradient of the chained MLP (conv -> BN(batch stats) -> ReLU
    [-> max over 64 neighbours]) against the same modules in float64 torch.
    relu=False: nothing but the pooled arg-max decides where a gradient goes (and the reference is given our arg-max), so
    every gradient is held to 2e-4 of its peak, element by element.  relu=True: the ReLU mask of an activation within fp32
    rounding of zero differs between fp32 and float64 (about one element per million), which moves O(1) gradient at those
    positions and, through the batch means of the BatchNorm backward, O(1/positions) gradient in the whole channel (these
    test chains have only ~2 500 positions per channel); gradients are then held to 2e-3 after discarding the 1 % largest
    deviations, and to 2e-2 in relative L2."""
    from regnet_for_3d_grasping_b200 import conv_train as ct
    ndim = len(shape) - 2
    ours, ref = _mlp_pair(shape[1], widths, ndim, relu=relu)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(shape, generator=g)
    xo = x.cuda().requires_grad_(True)
    xr = x.double().requires_grad_(True)
    assert ct.chain_supported(ours, xo)
    arg = None
    if pooled:
        # the reference routes the pooled gradient to the position OUR forward picked (near-ties inside a row)
        import copy
        with torch.no_grad():
            arg = copy.deepcopy(ours)(xo.detach()).argmax(dim=3, keepdim=True).cpu()
    yo = ours.forward_max_over_neighbours(xo) if pooled else ours(xo)
    monkeypatch.setenv("REGNET_TRAIN_TORCH", "1")
    yr = ref(xr)
    if pooled:
        assert (yr.detach().gather(3, arg).squeeze(3) - yr.detach().max(dim=3)[0]).abs().max() < 1e-4 * yr.detach().abs().max()
        yr = yr.gather(3, arg).squeeze(3)
    monkeypatch.delenv("REGNET_TRAIN_TORCH")
    dy = torch.randn(yr.shape, generator=g)
    yo.backward(dy.cuda())
    yr.backward(dy.double())
    torch.cuda.synchronize()
    assert _rel(yo, yr.detach()) < 1e-4

    # A gradient that is analytically zero (e.g. the BatchNorm shift of a block without ReLU in front of another
    # BatchNorm) is compared against the largest gradient of its kind in the chain, not against itself.
    floor = {"dW": max(float(b.conv.weight.grad.abs().max()) for b in ref),
             "dgamma": max(float(b.bn.weight.grad.abs().max()) for b in ref),
             "dbeta": max(float(b.bn.bias.grad.abs().max()) for b in ref), "dx": 0.0}

    def grad_ok(got, want, kind, what):
        got, want = got.double().cpu(), want.cpu()
        scale = max(float(want.abs().max()), floor[kind], 1e-30)
        err = (got - want).abs().flatten()
        if relu:
            k = int(0.01 * err.numel())
            worst = float(err.kthvalue(err.numel() - k)[0]) if k > 0 else float(err.max())
            assert worst < 2e-3 * scale, (what, worst, scale)
            assert float((got - want).norm()) < 2e-2 * max(float(want.norm()), scale), what
        else:
            assert float(err.max()) < 2e-4 * scale, (what, float(err.max()), scale)
    grad_ok(xo.grad, xr.grad, "dx", "dx")
    for i, (bo, br) in enumerate(zip(ours, ref)):
        grad_ok(bo.conv.weight.grad, br.conv.weight.grad, "dW", f"dW{i}")
        grad_ok(bo.bn.weight.grad, br.bn.weight.grad, "dgamma", f"dgamma{i}")
        grad_ok(bo.bn.bias.grad, br.bn.bias.grad, "dbeta", f"dbeta{i}")
        assert _rel(bo.bn.running_mean, br.bn.running_mean) < 1e-4
        assert _rel(bo.bn.running_var, br.bn.running_var) < 1e-4
        assert int(bo.bn.num_batches_tracked) == 1


def test_mlp_chain_dropout_is_a_consistent_mask(lib_path):
    """Dropout inside the chain: the kept fraction is 1 - p, kept values are scaled by 1 / (1 - p), and the backward uses the
    same mask (the gradient w.r.t. a dropped output element does not reach the input)."""
    ours, _ = _mlp_pair(32, (64,), 1, dropout=0.5)
    x = torch.randn(2, 32, 4096, device="cuda")
    torch.manual_seed(3)
    y = ours(x)
    # reference values without dropout, from the same batch statistics
    import copy
    nod = copy.deepcopy(ours)
    nod.dropout_prob = 0.0
    y0 = nod(x)
    pos = y0 > 1e-6
    kept = (y != 0) & pos
    frac = kept.sum().item() / pos.sum().item()
    assert abs(frac - 0.5) < 0.01, frac
    assert torch.allclose(y[kept], 2.0 * y0[kept], rtol=1e-5, atol=1e-6)
    # same seed -> same mask
    torch.manual_seed(3)
    assert torch.equal(ours(x) != 0, y != 0)
    # backward: gradient only flows through kept elements
    xg = x.clone().requires_grad_(True)
    torch.manual_seed(3)
    yg = ours(xg)
    gsel = torch.zeros_like(yg)
    dropped = (~kept) & pos
    gsel[dropped] = 1.0
    yg.backward(gsel)
    # BatchNorm couples positions through the batch statistics, but with an all-dropped upstream gradient everything is zero
    assert float(xg.grad.abs().max()) == 0.0


def test_scorenet_train_step_uses_no_library_gemm(lib_path):
    """One ScoreNet training step (small cloud): no cuDNN / cuBLAS / cutlass kernel in the profile."""
    from torch.profiler import ProfilerActivity, profile
    from regnet_for_3d_grasping_b200 import synth, weights
    from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
    net = ScoreNetwork(training=True).cuda()
    net.load_state_dict(weights.random_scorenet_state(seed=0))
    net.train()
    pc = torch.from_numpy(synth.batch("table", [0, 1], 8192)).cuda()
    tgt = torch.from_numpy(synth.scores_like_dataset(7, 2, 8192)).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        _, _, loss = net(pc, tgt)
        loss.sum().backward()
        opt.step()
    step()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    bad = [n for n in names if any(t in n.lower() for t in ("cudnn", "cutlass", "cublas", "gemm", "nchwtonhwc", "sgemm", "xmma"))
           and "regnet" not in n]
    assert not bad, bad
    assert any("conv1x1_tc_kernel" in n for n in names) and any("wgrad_tc_kernel" in n for n in names)


def _clone_module(mod):
    import copy
    return copy.deepcopy(mod)


def _grads(mod):
    return {k: p.grad.clone() for k, p in mod.named_parameters() if p.grad is not None}


def _assert_same_grads(a, b, tol=2e-5):
    """Same gradients up to fp32 summation order.  A gradient that is analytically zero (the BatchNorm shift of a block in
    front of another BatchNorm) is pure rounding noise: it is measured against the largest gradient of the set."""
    assert a.keys() == b.keys()
    top = max(float(v.abs().max()) for v in b.values())
    for k in a:
        scale = max(float(b[k].abs().max()), 1e-2 * top, 1e-30)
        assert float((a[k] - b[k]).abs().max()) <= tol * scale, k


def test_fused_operand_producers_match_the_unfused_modules(lib_path, monkeypatch):
    """PointNetSAModule / PointnetFPModule in train mode: grouping (or interpolation) + concat written directly as operand
    planes and the strided scatter-add backward, against the op-by-op module path (group_points / feature_interpolate /
    torch.cat / split) on the same engine: same planes, hence the same values; gradients up to the scatter-add order."""
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.modules import PointNetSAModule, PointnetFPModule
    torch.manual_seed(0)
    pts = torch.from_numpy(synth.batch("table", [3, 4], 2048)).cuda()
    xyz = pts[:, :, :3].permute(0, 2, 1)                      # strided views, as PointNet2Seg passes them
    feat0 = torch.randn(2, 2048, 19, device="cuda").permute(0, 2, 1)
    sa = PointNetSAModule(19, (32, 64), 256, 0.1, 64, use_xyz=True).cuda().train()
    fp = PointnetFPModule(64 + 19, (48, 40), 3).cuda().train()
    res = []
    for unfused in ("0", "1"):
        monkeypatch.setenv("REGNET_TRAIN_UNFUSED_OPERANDS", unfused)
        sa_i, fp_i = _clone_module(sa), _clone_module(fp)
        f = feat0.clone().requires_grad_(True)
        new_xyz, new_feat = sa_i(xyz, f)
        out = fp_i(xyz, new_xyz, f, new_feat)
        g = torch.Generator(device="cuda").manual_seed(5)
        (out * torch.randn(out.shape, device="cuda", generator=g)).sum().backward()
        res.append((new_feat.detach(), out.detach(), f.grad.clone(), _grads(sa_i), _grads(fp_i)))
    (nf0, o0, df0, gs0, gf0), (nf1, o1, df1, gs1, gf1) = res
    # (the batch moments are accumulated with fp64 atomics in arbitrary order: statistics may differ in the last bit)
    assert float((nf0 - nf1).abs().max()) <= 2e-6 * float(nf1.abs().max())
    assert float((o0 - o1).abs().max()) <= 2e-6 * float(o1.abs().max())
    assert float((df0 - df1).abs().max()) <= 2e-5 * float(df1.abs().max())
    _assert_same_grads(gs0, gs1)
    _assert_same_grads(gf0, gf1)


def test_bn_backward_reduction_fused_into_dgrad(lib_path, monkeypatch):
    """The reduction pass of a block's BatchNorm backward fused into the dgrad epilogue of the block above (with and
    without dropout) against the separate reduction kernel: same sums up to fp32 summation order."""
    res = []
    for drop in (0.0, 0.5):
        pair = []
        for fused in ("1", "0"):
            monkeypatch.setenv("REGNET_TRAIN_FUSED_BNREDUCE", fused)
            ours, _ = _mlp_pair(40, (64, 96, 32), 1, dropout=drop, seed=3)
            x = torch.randn(3, 40, 2056, generator=torch.Generator().manual_seed(2)).cuda().requires_grad_(True)
            torch.manual_seed(9)
            y = ours(x)
            y.backward(torch.randn(y.shape, generator=torch.Generator().manual_seed(4)).cuda())
            pair.append((y.detach(), x.grad.clone(), _grads(ours)))
        (y0, dx0, g0), (y1, dx1, g1) = pair
        assert float((y0 - y1).abs().max()) <= 2e-6 * float(y1.abs().max())
        assert float((dx0 - dx1).abs().max()) <= 2e-5 * float(dx1.abs().max())
        _assert_same_grads(g0, g1)


def test_linear_first_bodies_match_the_grouped_ones(lib_path, monkeypatch):
    """First convolution applied per source point (SA levels with >= 64 input channels, the last FP module) against the
    same bodies with the convolution on the grouped / interpolated positions: linear operations commute, so values and
    gradients agree to the arithmetic's accuracy (two different split-bf16 evaluation orders: 1e-5)."""
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.modules import PointNetSAModule, PointnetFPModule
    torch.manual_seed(1)
    pts = torch.from_numpy(synth.batch("table", [5, 6], 2048)).cuda()
    xyz = pts[:, :, :3].permute(0, 2, 1)
    rgb = pts[:, :, 3:6].permute(0, 2, 1)
    feat0 = torch.randn(2, 64, 2048, device="cuda")
    sa = PointNetSAModule(64, (96, 128), 256, 0.1, 64, use_xyz=True).cuda().train()
    fp = PointnetFPModule(128 + 3, (80, 48), 3).cuda().train()
    res = []
    for linear in ("1", "0"):
        monkeypatch.setenv("REGNET_TRAIN_LINEAR_FIRST", linear)
        sa_i, fp_i = _clone_module(sa), _clone_module(fp)
        f = feat0.clone().requires_grad_(True)
        new_xyz, new_feat = sa_i(xyz, f)
        out = fp_i(xyz, new_xyz, rgb, new_feat)                     # dense = rgb: 3 channels, no gradient
        g = torch.Generator(device="cuda").manual_seed(5)
        (out * torch.randn(out.shape, device="cuda", generator=g)).sum().backward()
        res.append((new_feat.detach(), out.detach(), f.grad.clone(), _grads(sa_i), _grads(fp_i)))
    (nf0, o0, df0, gs0, gf0), (nf1, o1, df1, gs1, gf1) = res
    assert float((nf0 - nf1).abs().max()) <= 5e-5 * float(nf1.abs().max())
    assert float((o0 - o1).abs().max()) <= 1e-4 * float(o1.abs().max())
    # gradients pass through ReLU / arg-max decisions that differ for values within 1e-5 of a boundary (the two bodies
    # evaluate the first convolution in different orders): robust comparison
    def close(a, b, what, floor=0.0):
        err = (a - b).abs().flatten()
        k = max(1, int(0.01 * err.numel()))
        scale = max(float(b.abs().max()), floor, 1e-30)      # analytically-zero gradients (BN shifts in front of a BN) are noise
        worst = float(err.kthvalue(err.numel() - k + 1)[0])
        # (scripts/linear_first_debug.py: both bodies sit at the same distance, to 4 digits, from a float64 evaluation.)
        # Per-channel vectors (80 numbers, each a sum over every position of mask-gated terms) have no 1 % of outliers to
        # drop: a handful of flipped masks moves single entries, so they are held to the L2 bound only.
        assert a.numel() < 1000 or worst <= 1e-2 * scale, (what, worst, scale)
        assert float((a - b).norm()) <= 2e-2 * max(float(b.norm()), scale * b.numel() ** 0.5), what
    close(df0, df1, "dfeature")
    for grads0, grads1, name in ((gs0, gs1, "sa."), (gf0, gf1, "fp.")):
        top = max(float(v.abs().max()) for v in grads1.values())
        for k in grads0:
            close(grads0[k], grads1[k], name + k, floor=0.1 * top)


def test_level0_recompute_body_matches_the_grouped_one(lib_path, monkeypatch):
    """Set-abstraction level 0 without its pre-activation (moments from the 27 input sums, activation recomputed, backward from
    7 sums per channel) against the grouped body that materialises Z0: same function; the recomputed Z0 is an fp32 FMA
    chain where the grouped body runs a split-bf16 GEMM, so values agree to ~1e-5 and gradients up to ReLU / arg-max flips."""
    from regnet_for_3d_grasping_b200 import synth
    from regnet_for_3d_grasping_b200.modules import PointNetSAModule
    torch.manual_seed(2)
    pts = torch.from_numpy(synth.batch("table", [7, 8, 9], 4096)).cuda()
    xyz = pts[:, :, :3].permute(0, 2, 1)
    rgb = pts[:, :, 3:6].permute(0, 2, 1)
    sa = PointNetSAModule(3, (128, 128, 256), 512, 0.06, 64, use_xyz=True).cuda().train()
    with torch.no_grad():
        for m in sa.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("REGNET_TRAIN_SA0_RECOMPUTE", flag)
        sa_i = _clone_module(sa)
        new_xyz, new_feat = sa_i(xyz, rgb)
        if flag == "1":
            from regnet_for_3d_grasping_b200 import pn2_ext
            geometry = (new_xyz.detach(), pn2_ext.ball_query(xyz, new_xyz, 0.06, 64)[0])
        g = torch.Generator(device="cuda").manual_seed(5)
        (new_feat * torch.randn(new_feat.shape, device="cuda", generator=g)).sum().backward()
        bufs = {k: v.clone() for k, v in sa_i.named_buffers()}
        res.append((new_feat.detach(), _grads(sa_i), bufs))
    (y0, g0, b0), (y1, g1, b1) = res
    assert float((y0 - y1).abs().max()) <= 5e-5 * float(y1.abs().max())
    for k in b0:                                      # running statistics of every block, the closed-form ones included
        assert torch.allclose(b0[k].float(), b1[k].float(), rtol=1e-4, atol=1e-6), k
    top = max(float(v.abs().max()) for v in g1.values())
    for k in g0:
        a, b = g0[k], g1[k]
        err = (a - b).abs().flatten()
        kk = max(1, int(0.01 * err.numel()))
        scale = max(float(b.abs().max()), 0.1 * top)
        assert a.numel() < 1000 or float(err.kthvalue(err.numel() - kk + 1)[0]) <= 1e-2 * scale, k
        assert float((a - b).norm()) <= 2e-2 * max(float(b.norm()), scale * b.numel() ** 0.5), k
    # and the first block's own parameters against autograd through a plain fp64 evaluation of the same body
    monkeypatch.setenv("REGNET_TRAIN_TORCH", "1")
    sa_t = _clone_module(sa).double()
    _, ft = sa_t(xyz.double(), rgb.double(), geometry=(geometry[0].double(), geometry[1]))
    g = torch.Generator(device="cuda").manual_seed(5)
    (ft * torch.randn(ft.shape, device="cuda", generator=g).double()).sum().backward()
    gt = _grads(sa_t)
    for k in ("mlp.0.conv.weight", "mlp.0.bn.weight", "mlp.0.bn.bias"):
        rel = float((g0[k].double() - gt[k]).norm() / gt[k].norm())
        rel_grouped = float((g1[k].double() - gt[k]).norm() / gt[k].norm())
        assert rel <= max(2.0 * rel_grouped, 2e-3), (k, rel, rel_grouped)
