"""Training-path shared MLP on this repo's tensor-core engine (csrc/conv_train.cu + csrc/train_ops.cu).

Mirrors what torch executes for the reference's SharedMLP in train mode
(multi_model/utils/pn2_utils/nn/modules/mlp.py:95-106 over nn/modules/conv.py:24-36,64-76: 1x1 conv without bias ->
BatchNorm with batch statistics -> ReLU [-> dropout], and for set-abstraction modules the max over the 64 neighbours of
modules.py:245), forward and backward, as ONE autograd function per MLP:

  forward   per block: Z = W x on tcgen05 (operands as bf16 hi/lo planes, the batch moments of Z accumulated in the GEMM
            epilogue) -> statistics -> y = dropout(relu(bn(Z))) written directly as the next block's planes
  backward  per block: (dy, Z) -> dZ planes (+ dgamma, dbeta) -> dW = dZ x^T (split-K wgrad), dx = W^T dZ (dgrad)

No cuDNN / cuBLAS kernel runs on this path.  `passes` selects the arithmetic: 3 = split-bf16 (fp32 parity, default),
1 = plain bf16 (REGNET_TRAIN_PASSES=1)."""
import os

import torch

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return _lib.current_stream_ptr()


def default_passes():
    return 1 if os.environ.get("REGNET_TRAIN_PASSES", "3") == "1" else 3


def enabled():
    """REGNET_TRAIN_TORCH=1 keeps torch's modules; REGNET_TRAIN_CONV_TORCH=1 keeps only the convolutions on torch."""
    return os.environ.get("REGNET_TRAIN_TORCH", "0") != "1" and os.environ.get("REGNET_TRAIN_CONV_TORCH", "0") != "1"


def _round8(n):
    return (n + 7) // 8 * 8


# ---- thin wrappers over the C ABI ------------------------------------------------------------------------------------
def split_planes(x):
    """fp32 tensor -> (hi, lo) bf16 tensors of the same shape with x ~= hi + lo."""
    x = x.contiguous()
    hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    if x.numel():
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().regnet_split_planes(_p(x), x.numel(), _p(hi), _p(lo), _stream()))
    return hi, lo


def split_weight(w2d, transpose=False):
    """(rows, cols) fp32 -> planes (rows, ld) or, transposed, (cols, ld); ld = columns rounded up to 8, zero padded."""
    w2d = w2d.contiguous()
    rows, cols = w2d.shape
    out_rows, out_cols = (cols, rows) if transpose else (rows, cols)
    ld = _round8(out_cols)
    hi = torch.empty(out_rows, ld, dtype=torch.bfloat16, device=w2d.device)
    lo = torch.empty(out_rows, ld, dtype=torch.bfloat16, device=w2d.device)
    with torch.cuda.device(w2d.device):
        _lib.check(_lib.load().regnet_split_weight(_p(w2d), rows, cols, int(transpose), ld, _p(hi), _p(lo), _stream()))
    return hi, lo


def conv1x1(x_hi, x_lo, a_hi, a_lo, rows, K, want_moments=False, passes=3):
    """out[b, r, l] = sum_k A[r, k] x[b, k, l] for x planes (B, K, L) -> fp32 (B, rows, L) [, moments (rows, 2) fp64]."""
    B, Kx, L = x_hi.shape
    assert Kx == K and a_hi.shape[0] == rows and a_hi.shape[1] >= K
    out = torch.empty(B, rows, L, dtype=torch.float32, device=x_hi.device)
    moments = torch.empty(rows, 2, dtype=torch.float64, device=x_hi.device) if want_moments else None
    with torch.cuda.device(x_hi.device):
        _lib.check(_lib.load().regnet_conv1x1_train(_p(x_hi), _p(x_lo), B, K, L, _p(a_hi), _p(a_lo), rows, a_hi.shape[1],
                                                    _p(out), _p(moments), passes, _stream()))
    return (out, moments) if want_moments else out


def wgrad(g_hi, g_lo, x_hi, x_lo, passes=3):
    """dW[co, ci] = sum_{b,l} g[b, co, l] x[b, ci, l] for planes (B, Co, L), (B, Ci, L) -> fp32 (Co, Ci)."""
    B, Co, L = g_hi.shape
    Ci = x_hi.shape[1]
    lib = _lib.load()
    dW = torch.empty(Co, Ci, dtype=torch.float32, device=g_hi.device)
    with torch.cuda.device(g_hi.device):
        nbytes = int(lib.regnet_conv1x1_wgrad_workspace_bytes(B, Co, Ci, L))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=g_hi.device)
        _lib.check(lib.regnet_conv1x1_train_wgrad(_p(g_hi), _p(g_lo), _p(x_hi), _p(x_lo), B, Co, Ci, L, _p(dW), _p(ws),
                                                  nbytes, passes, _stream()))
    return dW


def _bn_ws(lib, B, C, L, device):
    nbytes = int(lib.regnet_bn_workspace_bytes(B, C, L))
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


# ---- the chained MLP ---------------------------------------------------------------------------------------------------
class _Spec:
    """Static description of one chain call (not a tensor: passed to the autograd function as a plain object)."""

    def __init__(self, blocks, pooled, dropout_p, passes, seeds):
        self.bns = [b.bn for b in blocks]
        self.relu = [b.relu is not None for b in blocks]
        self.pooled = pooled
        self.dropout_p = float(dropout_p)
        self.passes = passes
        self.seeds = seeds          # one dropout seed per block (unused when dropout_p == 0)


def _chain_forward(hi, lo, B, L, spec, weights, gammas, betas, first=None):
    """Blocks of one shared MLP over input planes (B, C0, L): -> (out fp32, saved tensors [hi, lo, z, stats] per block, arg).
    first = (z0, moments0): the first block's convolution output and batch moments were produced elsewhere (the
    linear-first bodies below); hi / lo are then unused placeholders."""
    lib = _lib.load()
    dev = weights[0].device
    n = len(spec.bns)
    saved, out, arg = [], None, None
    for i in range(n):
        w = weights[i]
        cout, cin = w.shape[0], w.shape[1]
        if i == 0 and first is not None:
            z, moments = first
            hi = lo = torch.empty(0, dtype=torch.bfloat16, device=dev)
        else:
            a_hi, a_lo = split_weight(w.reshape(cout, cin))
            z, moments = conv1x1(hi, lo, a_hi, a_lo, cout, cin, want_moments=True, passes=spec.passes)
        bn = spec.bns[i]
        stats = torch.empty(4, cout, dtype=torch.float32, device=dev)   # mean, invstd, scale, shift
        _lib.check(lib.regnet_bn_finalize_moments(
            _p(moments), cout, float(B) * float(L), _p(gammas[i]), _p(betas[i]), float(bn.eps), float(bn.momentum),
            _p(bn.running_mean), _p(bn.running_var), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), _stream()))
        saved += [hi, lo, z, stats]
        last = i == n - 1
        if last and spec.pooled:
            M = L // 64
            out = torch.empty(B, cout, M, dtype=torch.float32, device=dev)
            arg = torch.empty(B, cout, M, dtype=torch.uint8, device=dev)
            _lib.check(lib.regnet_bn_apply_max64(_p(z), B, cout, M, _p(stats[2]), _p(stats[3]), int(spec.relu[i]),
                                                 _p(out), _p(arg), _stream()))
        elif last:
            out = torch.empty(B, cout, L, dtype=torch.float32, device=dev)
            _lib.check(lib.regnet_bn_apply_ex(_p(z), B, cout, L, _p(stats[2]), _p(stats[3]), int(spec.relu[i]),
                                              spec.dropout_p, spec.seeds[i], _p(out), None, None, _stream()))
        else:
            hi = torch.empty(B, cout, L, dtype=torch.bfloat16, device=dev)
            lo = torch.empty(B, cout, L, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.regnet_bn_apply_ex(_p(z), B, cout, L, _p(stats[2]), _p(stats[3]), int(spec.relu[i]),
                                              spec.dropout_p, spec.seeds[i], None, _p(hi), _p(lo), _stream()))
    return out, saved, arg


def _chain_backward(dout, B, L, spec, weights, saved, arg, need_w, need_dx, first_linear=False):
    """-> (dx fp32 (B, C0, L) or None, dW list, dgamma list, dbeta list).
    first_linear: block 0's convolution lives outside the chain; its BatchNorm backward then writes the gradient w.r.t.
    the convolution OUTPUT as fp32 and that tensor is returned in place of dx (dW[0] stays None)."""
    lib = _lib.load()
    dev = dout.device
    n = len(spec.bns)
    dws, dgs, dbs = [None] * n, [None] * n, [None] * n
    dy = dout.contiguous()
    # The reduction pass of a block's BatchNorm backward can ride in the dgrad epilogue of the block above
    # (REGNET_TRAIN_FUSED_BNREDUCE=1).  Measured on B200 it is a loss: the per-thread 128-byte loads of z stretch the
    # epilogue (dgrad 7.0 -> 9.5 ms per step) by more than the separate HBM-roofline pass costs (3.3 ms), so it is off.
    fuse = os.environ.get("REGNET_TRAIN_FUSED_BNREDUCE", "0") == "1"
    sums = None                     # (C, 2) fp64 from the dgrad epilogue of the block above, when fused
    for i in range(n - 1, -1, -1):
        hi, lo, z, stats = saved[4 * i:4 * i + 4]
        w = weights[i]
        cout, cin = w.shape[0], w.shape[1]
        g_hi = torch.empty(B, cout, L, dtype=torch.bfloat16, device=dev)
        g_lo = torch.empty(B, cout, L, dtype=torch.bfloat16, device=dev)
        dgamma = torch.empty(cout, dtype=torch.float32, device=dev)
        dbeta = torch.empty(cout, dtype=torch.float32, device=dev)
        ws, nbytes = _bn_ws(lib, B, cout, L, dev)
        if i == 0 and first_linear:
            del g_hi, g_lo
            dz0 = torch.empty(B, cout, L, dtype=torch.float32, device=dev)
            if n == 1 and spec.pooled:
                _lib.check(lib.regnet_bn_max64_backward_ex(
                    _p(dy), _p(arg), _p(z), B, cout, L // 64, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
                    int(spec.relu[i]), _p(dz0), None, None, _p(dgamma), _p(dbeta), _p(ws), nbytes, _stream()))
            elif sums is not None:
                _lib.check(lib.regnet_bn_backward_from_sums(
                    _p(dy), _p(z), B, cout, L, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), int(spec.relu[i]),
                    spec.dropout_p, spec.seeds[i], _p(sums), _p(dz0), None, None, _p(dgamma), _p(dbeta), _p(ws), nbytes,
                    _stream()))
            else:
                _lib.check(lib.regnet_bn_backward_ex(
                    _p(dy), _p(z), B, cout, L, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), int(spec.relu[i]),
                    spec.dropout_p, spec.seeds[i], _p(dz0), None, None, _p(dgamma), _p(dbeta), _p(ws), nbytes, _stream()))
            dgs[i], dbs[i] = dgamma, dbeta
            return dz0, dws, dgs, dbs
        if i == n - 1 and spec.pooled:
            _lib.check(lib.regnet_bn_max64_backward_ex(
                _p(dy), _p(arg), _p(z), B, cout, L // 64, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
                int(spec.relu[i]), None, _p(g_hi), _p(g_lo), _p(dgamma), _p(dbeta), _p(ws), nbytes, _stream()))
        elif sums is not None:
            _lib.check(lib.regnet_bn_backward_from_sums(
                _p(dy), _p(z), B, cout, L, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
                int(spec.relu[i]), spec.dropout_p, spec.seeds[i], _p(sums), None, _p(g_hi), _p(g_lo), _p(dgamma),
                _p(dbeta), _p(ws), nbytes, _stream()))
        else:
            _lib.check(lib.regnet_bn_backward_ex(
                _p(dy), _p(z), B, cout, L, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]),
                int(spec.relu[i]), spec.dropout_p, spec.seeds[i], None, _p(g_hi), _p(g_lo), _p(dgamma), _p(dbeta),
                _p(ws), nbytes, _stream()))
        dgs[i], dbs[i] = dgamma, dbeta
        sums = None
        if need_w[i]:
            dws[i] = wgrad(g_hi, g_lo, hi, lo, passes=spec.passes).view(w.shape)
        if i > 0 or need_dx:
            t_hi, t_lo = split_weight(w.reshape(cout, cin), transpose=True)
            if i > 0 and fuse:
                # dgrad + the reduction pass of block i-1's BatchNorm backward in one kernel
                zp, sp = saved[4 * (i - 1) + 2], saved[4 * (i - 1) + 3]
                dy = torch.empty(B, cin, L, dtype=torch.float32, device=dev)
                sums = torch.empty(cin, 2, dtype=torch.float64, device=dev)
                _lib.check(lib.regnet_conv1x1_train_dgrad_bnreduce(
                    _p(g_hi), _p(g_lo), B, cout, L, _p(t_hi), _p(t_lo), cin, t_hi.shape[1], _p(dy), _p(zp), _p(sp[0]),
                    _p(sp[1]), _p(sp[2]), _p(sp[3]), int(spec.relu[i - 1]), spec.dropout_p, spec.seeds[i - 1],
                    _p(sums), spec.passes, _stream()))
            else:
                dy = conv1x1(g_hi, g_lo, t_hi, t_lo, cin, cout, passes=spec.passes)
        else:
            dy = None
        del g_hi, g_lo
    return dy, dws, dgs, dbs


class _MLPChainTrain(torch.autograd.Function):
    """Shared MLP on an fp32 input tensor (B, C0, ...)."""

    @staticmethod
    def forward(ctx, x, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        shape = x.shape
        B, C0 = shape[0], shape[1]
        L = x.numel() // max(B * C0, 1)
        with torch.cuda.device(x.device):
            hi, lo = split_planes(x.contiguous().view(B, C0, L))
            out, saved, arg = _chain_forward(hi, lo, B, L, spec, weights, gammas, betas)
        ctx.spec, ctx.n, ctx.in_shape, ctx.L = spec, n, shape, L
        ctx.save_for_backward(*(list(weights) + saved + ([arg] if arg is not None else [])))
        if spec.pooled:
            return out.view(B, out.shape[1], shape[2])
        return out.view((B, out.shape[1]) + tuple(shape[2:]))

    @staticmethod
    def backward(ctx, dout):
        spec, n, L = ctx.spec, ctx.n, ctx.L
        tensors = ctx.saved_tensors
        weights, saved = tensors[:n], tensors[n:n + 4 * n]
        arg = tensors[n + 4 * n] if spec.pooled else None
        B = ctx.in_shape[0]
        with torch.cuda.device(dout.device):
            dx, dws, dgs, dbs = _chain_backward(dout, B, L, spec, weights, saved, arg,
                                                [ctx.needs_input_grad[2 + i] for i in range(n)], ctx.needs_input_grad[0])
        dx = dx.view(ctx.in_shape) if dx is not None else None
        return (dx, None) + tuple(dws) + tuple(dgs) + tuple(dbs)


class _SAGroupedChainTrain(torch.autograd.Function):
    """Set-abstraction body in train mode: QueryGrouper (pn2_utils/modules.py:39-56: group xyz and features by the ball-query
    index, centre the coordinates, concat [xyz_rel | feature]) written directly as operand planes, then the pooled shared MLP.
    `feature` (B, C, N) is the only differentiable input; xyz / new_xyz / index ride in the spec."""

    @staticmethod
    def forward(ctx, feature, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        xyz, new_xyz, index = spec.xyz, spec.new_xyz, spec.index
        B, _, N = xyz.shape
        M, K = index.shape[1], index.shape[2]
        C = feature.shape[1]
        L = M * K
        dev = xyz.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            hi = torch.empty(B, C + 3, L, dtype=torch.bfloat16, device=dev)
            lo = torch.empty(B, C + 3, L, dtype=torch.bfloat16, device=dev)
            nx = new_xyz.contiguous()
            _lib.check(lib.regnet_sa_group_planes(_p(xyz), xyz.stride(0), xyz.stride(1), xyz.stride(2), _p(nx), _p(feature),
                                                  feature.stride(0), feature.stride(1), feature.stride(2), _p(index), B, C, N,
                                                  M, K, _p(hi), _p(lo), _stream()))
            out, saved, arg = _chain_forward(hi, lo, B, L, spec, weights, gammas, betas)
        ctx.spec, ctx.n, ctx.dims = spec, n, (B, C, N, M, K)
        ctx.save_for_backward(*(list(weights) + saved + [arg]))
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, n = ctx.spec, ctx.n
        B, C, N, M, K = ctx.dims
        tensors = ctx.saved_tensors
        weights, saved, arg = tensors[:n], tensors[n:n + 4 * n], tensors[n + 4 * n]
        dfeat = None
        with torch.cuda.device(dout.device):
            dx, dws, dgs, dbs = _chain_backward(dout, B, M * K, spec, weights, saved, arg,
                                                [ctx.needs_input_grad[2 + i] for i in range(n)], ctx.needs_input_grad[0])
            if dx is not None:
                # scatter-add of the feature channels (3 ..) of the operand gradient, read in place
                dfeat = torch.empty(B, C, N, dtype=torch.float32, device=dout.device)
                _lib.check(_lib.load().regnet_group_points_backward_strided(
                    _p(dx), (C + 3) * M * K, 3, _p(spec.index), B, C, N, M, K, _p(dfeat), _stream()))
        return (dfeat, None) + tuple(dws) + tuple(dgs) + tuple(dbs)


class _FPInterpChainTrain(torch.autograd.Function):
    """Feature-propagation body in train mode: FeatureInterpolator (modules.py:104-131: 3-NN weighted interpolation of the
    sparse features, concat [interpolated | dense]) written directly as operand planes, then the shared MLP."""

    @staticmethod
    def forward(ctx, sparse, dense, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        index, weight = spec.index, spec.weight
        B, C2, Ns = sparse.shape
        Nd = index.shape[1]
        C1 = 0 if dense is None else dense.shape[1]
        dev = sparse.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            hi = torch.empty(B, C2 + C1, Nd, dtype=torch.bfloat16, device=dev)
            lo = torch.empty(B, C2 + C1, Nd, dtype=torch.bfloat16, device=dev)
            ds = (0, 0, 0) if dense is None else dense.stride()
            _lib.check(lib.regnet_fp_interp_planes(_p(sparse), sparse.stride(0), sparse.stride(1), sparse.stride(2), _p(dense),
                                                   ds[0], ds[1], ds[2], _p(index), _p(weight), B, C2, C1, Ns, Nd, _p(hi),
                                                   _p(lo), _stream()))
            out, saved, _ = _chain_forward(hi, lo, B, Nd, spec, weights, gammas, betas)
        ctx.spec, ctx.n, ctx.dims = spec, n, (B, C2, C1, Ns, Nd)
        ctx.save_for_backward(*(list(weights) + saved))
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, n = ctx.spec, ctx.n
        B, C2, C1, Ns, Nd = ctx.dims
        tensors = ctx.saved_tensors
        weights, saved = tensors[:n], tensors[n:n + 4 * n]
        need_s, need_d = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and C1 > 0
        dsparse = ddense = None
        with torch.cuda.device(dout.device):
            dx, dws, dgs, dbs = _chain_backward(dout, B, Nd, spec, weights, saved, None,
                                                [ctx.needs_input_grad[3 + i] for i in range(n)], need_s or need_d)
            if need_s:
                dsparse = torch.empty(B, C2, Ns, dtype=torch.float32, device=dout.device)
                _lib.check(_lib.load().regnet_interpolate_backward_strided(
                    _p(dx), (C2 + C1) * Nd, 0, _p(spec.index), _p(spec.weight), B, C2, Ns, Nd, _p(dsparse), _stream()))
            if need_d:
                ddense = dx[:, C2:, :]
        return (dsparse, ddense, None) + tuple(dws) + tuple(dgs) + tuple(dbs)


class _SALinearChainTrain(torch.autograd.Function):
    """Set-abstraction body with the first convolution applied per SOURCE point: Y = W_f feature (a GEMM over the N points
    of the previous level instead of M * 64 grouped positions), Z0 = gather(Y) + W_x (xyz - centre) with its batch moments
    (regnet_sa_gather_linear), then the rest of the pooled MLP; backward = scatter-add of dZ0 (+ the tiny W_x reduction),
    wgrad / dgrad over N points.  Same function as _SAGroupedChainTrain (linear operations commute with the grouping)."""

    @staticmethod
    def forward(ctx, feature, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        xyz, new_xyz, index = spec.xyz, spec.new_xyz, spec.index
        B, _, N = xyz.shape
        M, K = index.shape[1], index.shape[2]
        C = feature.shape[1]
        L = M * K
        dev = xyz.device
        lib = _lib.load()
        w0 = weights[0].reshape(weights[0].shape[0], C + 3)
        C0 = w0.shape[0]
        with torch.cuda.device(dev):
            f_hi, f_lo = split_planes(feature.contiguous())
            wf_hi, wf_lo = split_weight(w0[:, 3:])
            y = conv1x1(f_hi, f_lo, wf_hi, wf_lo, C0, C, passes=spec.passes)                       # (B, C0, N)
            from . import pn2_ext
            xr = (pn2_ext.group_points_forward(xyz, index) - new_xyz.unsqueeze(-1)).view(B, 3, L)      # (B, 3, M*K), tiny
            wx = w0[:, :3].contiguous()
            z0 = torch.empty(B, C0, L, dtype=torch.float32, device=dev)
            moments = torch.empty(C0, 2, dtype=torch.float64, device=dev)
            _lib.check(lib.regnet_sa_gather_linear(_p(y), _p(index), _p(xr), _p(wx), 3, B, C0, N, M, K, _p(z0), _p(moments),
                                                   _stream()))
            del y
            out, saved, arg = _chain_forward(None, None, B, L, spec, weights, gammas, betas, first=(z0, moments))
        ctx.spec, ctx.n, ctx.dims = spec, n, (B, C, N, M, K, C0)
        ctx.save_for_backward(*(list(weights) + saved + [arg, f_hi, f_lo, xr]))
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, n = ctx.spec, ctx.n
        B, C, N, M, K, C0 = ctx.dims
        tensors = ctx.saved_tensors
        weights, saved = tensors[:n], tensors[n:n + 4 * n]
        arg, f_hi, f_lo, xr = tensors[n + 4 * n:n + 4 * n + 4]
        dev = dout.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            dz0, dws, dgs, dbs = _chain_backward(dout, B, M * K, spec, weights, saved, arg,
                                                 [ctx.needs_input_grad[2 + i] for i in range(n)], True, first_linear=True)
            dy = torch.empty(B, C0, N, dtype=torch.float32, device=dev)
            dwx_part = torch.empty(B, C0, 3, dtype=torch.float32, device=dev)
            _lib.check(lib.regnet_sa_scatter_linear(_p(dz0), _p(spec.index), _p(xr), B, C0, N, M, K, _p(dy), _p(dwx_part),
                                                    _stream()))
            del dz0
            g_hi, g_lo = split_planes(dy)
            dfeat = None
            w0 = weights[0].reshape(C0, C + 3)
            if ctx.needs_input_grad[2]:
                dwf = wgrad(g_hi, g_lo, f_hi, f_lo, passes=spec.passes)
                dws[0] = torch.cat([dwx_part.sum(dim=0), dwf], dim=1).view(weights[0].shape)
            if ctx.needs_input_grad[0]:
                t_hi, t_lo = split_weight(w0[:, 3:], transpose=True)
                dfeat = conv1x1(g_hi, g_lo, t_hi, t_lo, C, C0, passes=spec.passes)
        return (dfeat, None) + tuple(dws) + tuple(dgs) + tuple(dbs)


class _SubSpec:
    """Blocks [start:] of a chain spec (the same static fields, sliced)."""

    def __init__(self, spec, start):
        self.bns, self.relu, self.seeds = spec.bns[start:], spec.relu[start:], spec.seeds[start:]
        self.pooled, self.dropout_p, self.passes = spec.pooled, spec.dropout_p, spec.passes


class _SA0RecomputeChainTrain(torch.autograd.Function):
    """Set-abstraction level 0 (3 feature channels that need no gradient): the first block's pre-activation Z0 = W0 [xyz_rel |
    rgb] over the M * 64 grouped positions is never written.  Its batch moments follow from the 27 first and second sums of
    the 6 inputs, its activation is recomputed from the inputs as the next block's operand planes, and its whole backward
    (dW0, dgamma, dbeta) follows from 7 sums per channel taken in one pass over the incoming gradient (csrc/train_gather.cu,
    "set-abstraction level 0, first block")."""

    @staticmethod
    def forward(ctx, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        xyz, new_xyz, index, feature = spec.xyz, spec.new_xyz, spec.index, spec.feature
        B, _, N = xyz.shape
        M, K = index.shape[1], index.shape[2]
        L = M * K
        dev = xyz.device
        lib = _lib.load()
        C0 = weights[0].shape[0]
        w0 = weights[0].detach().reshape(C0, 6).contiguous()
        bn = spec.bns[0]
        with torch.cuda.device(dev):
            geo = (_p(xyz), xyz.stride(0), xyz.stride(1), xyz.stride(2), _p(new_xyz), _p(feature), feature.stride(0),
                   feature.stride(1), feature.stride(2), _p(index), B, N, M, K)
            sums = torch.empty(27, dtype=torch.float64, device=dev)
            moments = torch.empty(C0, 2, dtype=torch.float64, device=dev)
            _lib.check(lib.regnet_sa0_input_moments(*geo, _p(w0), C0, _p(sums), _p(moments), _stream()))
            stats = torch.empty(4, C0, dtype=torch.float32, device=dev)
            _lib.check(lib.regnet_bn_finalize_moments(
                _p(moments), C0, float(B) * float(L), _p(gammas[0]), _p(betas[0]), float(bn.eps), float(bn.momentum),
                _p(bn.running_mean), _p(bn.running_var), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), _stream()))
            hi = torch.empty(B, C0, L, dtype=torch.bfloat16, device=dev)
            lo = torch.empty(B, C0, L, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.regnet_sa0_apply_planes(*geo, _p(w0), _p(stats[2]), _p(stats[3]), C0, int(spec.relu[0]), _p(hi),
                                                   _p(lo), _stream()))
            out, saved, arg = _chain_forward(hi, lo, B, L, _SubSpec(spec, 1), weights[1:], gammas[1:], betas[1:])
        ctx.spec, ctx.n, ctx.dims = spec, n, (B, N, M, K, C0)
        ctx.save_for_backward(*(list(weights[1:]) + saved + [arg, w0, stats, sums]))
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, n = ctx.spec, ctx.n
        B, N, M, K, C0 = ctx.dims
        L = M * K
        tensors = ctx.saved_tensors
        weights, saved = tensors[:n - 1], tensors[n - 1:n - 1 + 4 * (n - 1)]
        arg, w0, stats, sums = tensors[n - 1 + 4 * (n - 1):]
        xyz, new_xyz, index, feature = spec.xyz, spec.new_xyz, spec.index, spec.feature
        dev = dout.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            dy0, dws, dgs, dbs = _chain_backward(dout, B, L, _SubSpec(spec, 1), weights, saved, arg,
                                                 [ctx.needs_input_grad[2 + i] for i in range(n - 1)], True)
            G = torch.empty(C0, 7, dtype=torch.float64, device=dev)
            _lib.check(lib.regnet_sa0_backward_sums(
                _p(xyz), xyz.stride(0), xyz.stride(1), xyz.stride(2), _p(new_xyz), _p(feature), feature.stride(0),
                feature.stride(1), feature.stride(2), _p(index), B, N, M, K, _p(dy0), _p(w0), _p(stats[2]), _p(stats[3]), C0,
                int(spec.relu[0]), _p(G), _stream()))
            del dy0
            dw0 = torch.empty(C0, 6, dtype=torch.float32, device=dev) if ctx.needs_input_grad[1] else None
            dgamma = torch.empty(C0, dtype=torch.float32, device=dev)
            dbeta = torch.empty(C0, dtype=torch.float32, device=dev)
            _lib.check(lib.regnet_sa0_backward_finalize(_p(G), _p(sums), _p(w0), _p(stats[1]), _p(stats[2]), C0,
                                                        float(B) * float(L), _p(dw0), _p(dgamma), _p(dbeta), _stream()))
            if dw0 is not None:
                dw0 = dw0.view(C0, 6, 1, 1)
        return (None, dw0) + tuple(dws) + (dgamma,) + tuple(dgs) + (dbeta,) + tuple(dbs)


class _FPLinearChainTrain(torch.autograd.Function):
    """Feature-propagation body with the first convolution applied before the interpolation: Ys = W_s sparse (a GEMM over the
    Ns sparse points instead of Nd dense ones), Z0 = interp(Ys) + W_d dense for a dense input of at most 4 channels that
    needs no gradient (the rgb skip of the last FP module), then the rest of the MLP."""

    @staticmethod
    def forward(ctx, sparse, spec, *params):
        n = len(spec.bns)
        weights, gammas, betas = params[:n], params[n:2 * n], params[2 * n:3 * n]
        index, weight, dense = spec.index, spec.weight, spec.dense
        B, C2, Ns = sparse.shape
        Nd = index.shape[1]
        nd = 0 if dense is None else dense.shape[1]
        dev = sparse.device
        lib = _lib.load()
        w0 = weights[0].reshape(weights[0].shape[0], C2 + nd)
        C0 = w0.shape[0]
        with torch.cuda.device(dev):
            s_hi, s_lo = split_planes(sparse.contiguous())
            ws_hi, ws_lo = split_weight(w0[:, :C2])
            ys = conv1x1(s_hi, s_lo, ws_hi, ws_lo, C0, C2, passes=spec.passes)                    # (B, C0, Ns)
            wd = w0[:, C2:].contiguous() if nd else None
            ds = (0, 0, 0) if dense is None else dense.stride()
            z0 = torch.empty(B, C0, Nd, dtype=torch.float32, device=dev)
            moments = torch.empty(C0, 2, dtype=torch.float64, device=dev)
            _lib.check(lib.regnet_fp_gather_linear(_p(ys), _p(index), _p(weight), _p(dense), ds[0], ds[1], ds[2], nd, _p(wd),
                                                   max(nd, 1), B, C0, Ns, Nd, _p(z0), _p(moments), _stream()))
            del ys
            out, saved, _ = _chain_forward(None, None, B, Nd, spec, weights, gammas, betas, first=(z0, moments))
        ctx.spec, ctx.n, ctx.dims = spec, n, (B, C2, nd, Ns, Nd, C0)
        ctx.save_for_backward(*(list(weights) + saved + [s_hi, s_lo]))
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, n = ctx.spec, ctx.n
        B, C2, nd, Ns, Nd, C0 = ctx.dims
        tensors = ctx.saved_tensors
        weights, saved = tensors[:n], tensors[n:n + 4 * n]
        s_hi, s_lo = tensors[n + 4 * n:n + 4 * n + 2]
        dev = dout.device
        lib = _lib.load()
        with torch.cuda.device(dev):
            dz0, dws, dgs, dbs = _chain_backward(dout, B, Nd, spec, weights, saved, None,
                                                 [ctx.needs_input_grad[2 + i] for i in range(n)], True, first_linear=True)
            from . import pn2_ext
            dys = pn2_ext.interpolate_backward(dz0, spec.index, spec.weight, Ns)                     # (B, C0, Ns)
            g_hi, g_lo = split_planes(dys)
            w0 = weights[0].reshape(C0, C2 + nd)
            dsparse = None
            if ctx.needs_input_grad[2]:
                dw_s = wgrad(g_hi, g_lo, s_hi, s_lo, passes=spec.passes)
                if nd:
                    dense = spec.dense
                    part = torch.empty(B, C0, nd, dtype=torch.float32, device=dev)
                    _lib.check(lib.regnet_fp_dense_wgrad(_p(dz0), _p(dense), dense.stride(0), dense.stride(1), dense.stride(2),
                                                         nd, B, C0, Nd, _p(part), _stream()))
                    dw_s = torch.cat([dw_s, part.sum(dim=0)], dim=1)
                dws[0] = dw_s.view(weights[0].shape)
            if ctx.needs_input_grad[0]:
                t_hi, t_lo = split_weight(w0[:, :C2], transpose=True)
                dsparse = conv1x1(g_hi, g_lo, t_hi, t_lo, C2, C0, passes=spec.passes)
        return (dsparse, None) + tuple(dws) + tuple(dgs) + tuple(dbs)


def chain_supported(mlp, x):
    """The chained path takes fp32 CUDA tensors whose per-entry position count is a multiple of 8, blocks of
    1x1 conv (no bias) + affine BatchNorm with running statistics."""
    if not (enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() in (3, 4) and x.numel() > 0):
        return False
    B, C = x.size(0), x.size(1)
    L = x.numel() // (B * C)
    if L % 8 != 0 or B * L <= 1 or B > 1023:
        return False
    if mlp.dropout_prob > 0.0 and mlp.ndim != 1:     # F.dropout2d drops whole channels: not the fused element mask
        return False
    for blk in mlp:
        bn, conv = blk.bn, blk.conv
        if bn is None or conv.bias is not None or not (bn.training and bn.affine and bn.track_running_stats):
            return False
        if bn.momentum is None or any(k != 1 for k in conv.kernel_size) or B * conv.out_channels > 65535:
            return False
    return True


def _spec_and_params(mlp, pooled):
    blocks = list(mlp)
    p = float(mlp.dropout_prob) if mlp.training else 0.0
    seeds = [int(s) for s in torch.randint(0, 2 ** 62, (len(blocks),))] if p > 0.0 else [0] * len(blocks)
    spec = _Spec(blocks, pooled, p, default_passes(), seeds)
    params = [b.conv.weight for b in blocks] + [b.bn.weight for b in blocks] + [b.bn.bias for b in blocks]
    return blocks, spec, params


def _count_batches(blocks):
    for b in blocks:
        if b.bn.num_batches_tracked is not None:
            b.bn.num_batches_tracked.add_(1)


def mlp_chain_train(mlp, x, pooled):
    """SharedMLP.forward (pooled=False) or torch.max(SharedMLP.forward(x), 3)[0] (pooled=True) in train mode."""
    blocks, spec, params = _spec_and_params(mlp, pooled)
    y = _MLPChainTrain.apply(x, spec, *params)
    _count_batches(blocks)
    return y


def _blocks_ok(mlp, B, L):
    if not enabled() or L % 8 != 0 or B * L <= 1 or B > 1023 or len(mlp) == 0:
        return False
    for blk in mlp:
        bn, conv = blk.bn, blk.conv
        if bn is None or conv.bias is not None or not (bn.training and bn.affine and bn.track_running_stats):
            return False
        if bn.momentum is None or any(k != 1 for k in conv.kernel_size) or B * conv.out_channels > 65535:
            return False
    return True


def _f32_cuda(*tensors):
    return all(t is not None and t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and t.numel() > 0 for t in tensors)


def _sa_linear_first(xyz, feature):
    return (os.environ.get("REGNET_TRAIN_LINEAR_FIRST", "1") != "0" and feature.size(1) >= 64 and xyz.size(2) <= 12288
            and feature.size(2) % 8 == 0)


def sa_grouped_supported(mlp, xyz, new_xyz, feature, index):
    """The fused set-abstraction body: 64 neighbours, a feature tensor, no dropout, source clouds whose scatter-add row
    fits in shared memory when the features need a gradient."""
    if not (mlp.training and mlp.ndim == 2 and mlp.dropout_prob == 0.0 and _f32_cuda(xyz, new_xyz, feature)):
        return False
    if index.dtype != torch.int64 or index.dim() != 3 or index.size(2) != 64 or not index.is_contiguous():
        return False
    if not _sa_linear_first(xyz, feature) and feature.requires_grad and xyz.size(2) > 12288:
        return False                                 # the grouped body's strided scatter-add keeps one row in shared memory
    if os.environ.get("REGNET_TRAIN_UNFUSED_OPERANDS", "0") == "1":
        return False
    return _blocks_ok(mlp, xyz.size(0), index.size(1) * 64)


def sa_grouped_chain_train(mlp, xyz, new_xyz, feature, index):
    """torch.max(mlp(cat([group(xyz) - new_xyz, group(feature)], 1)), 3)[0] -- modules.py:44-52 + :245 -- in train mode."""
    blocks, spec, params = _spec_and_params(mlp, True)
    spec.xyz, spec.new_xyz, spec.index = xyz.detach(), new_xyz.detach(), index
    # levels fed by a previous level (>= 64 feature channels over <= 12 288 source points): the first convolution runs per
    # source point, then a gather builds the grouped pre-activation (REGNET_TRAIN_LINEAR_FIRST=0 keeps it on the grouped
    # positions).  Level 0 (3 rgb channels over 25 600 points) keeps the grouped body: a 100 KB row per CTA leaves two CTAs
    # per SM and the shared-memory scatter-add of its 327 680 positions per row took 15 ms (measured).
    linear = _sa_linear_first(xyz, feature) and len(blocks) >= 1
    if (os.environ.get("REGNET_TRAIN_SA0_RECOMPUTE", "1") != "0" and feature.size(1) == 3 and not feature.requires_grad
            and len(blocks) >= 2 and spec.dropout_p == 0.0 and blocks[0].conv.weight.dim() == 4):
        # level 0: the first block's 2.5 GB pre-activation is never written (moments, activation and backward from its 6 inputs)
        spec.feature, spec.new_xyz = feature.detach(), new_xyz.detach().contiguous()
        y = _SA0RecomputeChainTrain.apply(spec, *params)
    else:
        fn = _SALinearChainTrain if linear else _SAGroupedChainTrain
        y = fn.apply(feature, spec, *params)
    _count_batches(blocks)
    return y


def fp_interp_supported(mlp, sparse_feature, dense_feature, index, weight):
    if not (mlp.training and mlp.ndim == 1 and _f32_cuda(sparse_feature) and (dense_feature is None or _f32_cuda(dense_feature))):
        return False
    if index.dtype != torch.int64 or index.dim() != 3 or index.size(2) != 3 or not (index.is_contiguous() and weight.is_contiguous()):
        return False
    if sparse_feature.size(2) > 12288 or index.size(1) % 8 != 0:
        return False
    if os.environ.get("REGNET_TRAIN_UNFUSED_OPERANDS", "0") == "1":
        return False
    return _blocks_ok(mlp, sparse_feature.size(0), index.size(1))


def fp_interp_chain_train(mlp, sparse_feature, dense_feature, index, weight):
    """mlp(cat([interpolate(sparse_feature, index, weight), dense_feature], 1)) -- modules.py:127-131, 508-509 -- in train mode."""
    blocks, spec, params = _spec_and_params(mlp, False)
    spec.index, spec.weight = index, weight.detach()
    small_dense = dense_feature is None or (dense_feature.size(1) <= 4 and not dense_feature.requires_grad)
    if (os.environ.get("REGNET_TRAIN_LINEAR_FIRST", "1") != "0" and small_dense and sparse_feature.size(2) % 8 == 0
            and sparse_feature.size(2) <= 12288):
        # the last FP module (512 interpolated channels + rgb): first convolution per SPARSE point, then interpolate
        spec.dense = None if dense_feature is None else dense_feature.detach()
        y = _FPLinearChainTrain.apply(sparse_feature, spec, *params)
    else:
        y = _FPInterpChainTrain.apply(sparse_feature, dense_feature, spec, *params)
    _count_batches(blocks)
    return y


class _Conv1x1Train(torch.autograd.Function):
    """A lone 1x1 convolution (optionally with bias) on the same engine: nn.Conv1d(k=1) of pointnet2.py:82 (conv_score)."""

    @staticmethod
    def forward(ctx, x, weight, bias, passes):
        B, C, L = x.shape
        cout = weight.shape[0]
        with torch.cuda.device(x.device):
            hi, lo = split_planes(x)
            a_hi, a_lo = split_weight(weight.reshape(cout, C))
            y = conv1x1(hi, lo, a_hi, a_lo, cout, C, passes=passes)
        if bias is not None:
            y += bias.view(1, cout, 1)
        ctx.save_for_backward(hi, lo, weight)
        ctx.has_bias = bias is not None
        ctx.passes = passes
        return y

    @staticmethod
    def backward(ctx, dy):
        hi, lo, weight = ctx.saved_tensors
        cout, C = weight.shape[0], weight.shape[1]
        dy = dy.contiguous()
        dx = dw = db = None
        with torch.cuda.device(dy.device):
            g_hi, g_lo = split_planes(dy)
            if ctx.needs_input_grad[1]:
                dw = wgrad(g_hi, g_lo, hi, lo, passes=ctx.passes).view(weight.shape)
            if ctx.needs_input_grad[0]:
                t_hi, t_lo = split_weight(weight.reshape(cout, C), transpose=True)
                dx = conv1x1(g_hi, g_lo, t_hi, t_lo, C, cout, passes=ctx.passes)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy.sum(dim=(0, 2))
        return dx, dw, db, None


def conv1x1_supported(conv, x):
    return (enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 3 and x.numel() > 0
            and x.size(2) % 8 == 0 and tuple(conv.kernel_size) == (1,) and tuple(conv.stride) == (1,)
            and conv.groups == 1)


def conv1x1_train(conv, x):
    """conv(x) for an nn.Conv1d with kernel size 1, through the tcgen05 engine (autograd-aware)."""
    return _Conv1x1Train.apply(x.contiguous(), conv.weight, conv.bias, default_passes())
