"""Training-mode pieces of the shared MLP on this repo's kernels (csrc/train_ops.cu), as autograd functions:

  bn_relu_train(x, bn, relu)   BatchNorm with batch statistics (+ ReLU) -- nn/modules/conv.py:24-36,64-76 in train mode
  max_over_neighbours(x)       torch.max(x, 3)[0] for K = 64 -- modules.py:245

Both fall back to torch (returning None / using torch.max) when the input is not a contiguous fp32 CUDA tensor of a
supported shape; REGNET_TRAIN_TORCH=1 forces the torch path (A/B measurements, parity tests)."""
import os

import torch

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return _lib.current_stream_ptr()


def fused_training_enabled():
    return os.environ.get("REGNET_TRAIN_TORCH", "0") != "1"


def _bn_shape(x):
    B, C = x.size(0), x.size(1)
    L = x.numel() // max(B * C, 1)
    return B, C, L


def bn_supported(x, bn):
    if not (fused_training_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and x.numel() > 0):
        return False
    if not (bn.training and bn.affine and bn.track_running_stats and bn.momentum is not None):
        return False
    B, C, L = _bn_shape(x)
    return L % 4 == 0 and B * C <= 65535 and B <= 1023 and B * L > 1


class _BnReluTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous()
        B, C, L = _bn_shape(x)
        lib = _lib.load()
        y = torch.empty_like(x)
        stats = torch.empty(4, C, dtype=torch.float32, device=x.device)   # save_mean, save_invstd, scale, shift
        with torch.cuda.device(x.device):
            nbytes = int(lib.regnet_bn_workspace_bytes(B, C, L))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(lib.regnet_bn_relu_train_forward(_p(x), B, C, L, _p(weight), _p(bias), float(eps), float(momentum),
                                                        int(relu), _p(running_mean), _p(running_var), _p(y), _p(stats[0]),
                                                        _p(stats[1]), _p(stats[2]), _p(stats[3]), _p(ws), nbytes, _stream()))
        ctx.save_for_backward(x, stats)
        ctx.relu = bool(relu)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        dy = dy.contiguous()
        B, C, L = _bn_shape(x)
        lib = _lib.load()
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = int(lib.regnet_bn_workspace_bytes(B, C, L))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(lib.regnet_bn_relu_train_backward(_p(dy), _p(x), B, C, L, _p(stats[0]), _p(stats[1]), _p(stats[2]),
                                                         _p(stats[3]), int(ctx.relu), _p(dx), _p(dgamma), _p(dbeta), _p(ws),
                                                         nbytes, _stream()))
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_relu_train(x, bn, relu):
    """y = [relu](bn(x)) with batch statistics; updates bn.running_* and num_batches_tracked like nn.BatchNorm*d."""
    y = _BnReluTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu)
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return y


class _BnReluMax64Train(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        x = x.contiguous()
        B, C, M = x.size(0), x.size(1), x.size(2)
        lib = _lib.load()
        out = torch.empty(B, C, M, dtype=torch.float32, device=x.device)
        arg = torch.empty(B, C, M, dtype=torch.uint8, device=x.device)
        stats = torch.empty(4, C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = int(lib.regnet_bn_workspace_bytes(B, C, M * 64))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(lib.regnet_bn_relu_max64_train_forward(
                _p(x), B, C, M, _p(weight), _p(bias), float(eps), float(momentum), int(relu), _p(running_mean),
                _p(running_var), _p(out), _p(arg), _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), _p(ws), nbytes,
                _stream()))
        ctx.save_for_backward(x, stats, arg)
        ctx.relu = bool(relu)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, stats, arg = ctx.saved_tensors
        dout = dout.contiguous()
        B, C, M = x.size(0), x.size(1), x.size(2)
        lib = _lib.load()
        dx = torch.empty_like(x)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            nbytes = int(lib.regnet_bn_workspace_bytes(B, C, M * 64))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(lib.regnet_bn_relu_max64_train_backward(
                _p(dout), _p(arg), _p(x), B, C, M, _p(stats[0]), _p(stats[1]), _p(stats[2]), _p(stats[3]), int(ctx.relu),
                _p(dx), _p(dgamma), _p(dbeta), _p(ws), nbytes, _stream()))
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_relu_max64_supported(x, bn):
    return x.dim() == 4 and x.size(3) == 64 and bn_supported(x, bn)


def bn_relu_max64_train(x, bn, relu):
    """max_k [relu](bn(x))[..., k] for x (B, C, M, 64) with batch statistics: the pooled block of a set-abstraction MLP
    (conv.py:64-76 + modules.py:245) without the (B, C, M, 64) activation."""
    y = _BnReluMax64Train.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, relu)
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    return y


class _MaxPool64(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        rows = x.numel() // 64
        out = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
        arg = torch.empty(x.shape[:-1], dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().regnet_maxpool64_forward(_p(x), rows, _p(out), _p(arg), _stream()))
        ctx.save_for_backward(arg)
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        dout = dout.contiguous()
        dx = torch.empty(tuple(arg.shape) + (64,), dtype=torch.float32, device=dout.device)
        with torch.cuda.device(dout.device):
            _lib.check(_lib.load().regnet_maxpool64_backward(_p(dout), _p(arg), arg.numel(), _p(dx), _stream()))
        return dx


def max_over_neighbours(x):
    """torch.max(x, 3)[0] for (B, C, M, K); this repo's kernel when K == 64 on a fp32 CUDA tensor."""
    if fused_training_enabled() and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.size(3) == 64 and x.numel() > 0:
        return _MaxPool64.apply(x)
    return torch.max(x, 3)[0]
