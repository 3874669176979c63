"""Parameter-holding layers with the reference's names and state-dict keys.

Mirrors multi_model/utils/pn2_utils/nn/modules/conv.py:6-83 (Conv1d / Conv2d = 1x1 conv without bias + BatchNorm +
ReLU), nn/modules/mlp.py:55-114 (SharedMLP) and nn/init.py:4-8 (init_bn), so that reference checkpoints load
(`<layer>.conv.weight`, `<layer>.bn.{weight,bias,running_mean,running_var,num_batches_tracked}`).
The forward here is the *training* path (torch conv + BN, autograd); in eval mode the fused plan in
scorenet.py consumes the same parameters folded to (W, scale, shift)."""
import os

import torch.nn.functional as F
from torch import nn

from . import conv_train, train_ops


def init_bn(module):
    if module.weight is not None:
        nn.init.ones_(module.weight)
    if module.bias is not None:
        nn.init.zeros_(module.bias)


class _ConvBnRelu(nn.Module):
    _conv_cls = None
    _bn_cls = None

    def __init__(self, in_channels, out_channels, kernel_size, relu=True, bn=True, bn_momentum=0.1, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.conv = self._conv_cls(in_channels, out_channels, kernel_size, bias=(not bn), **kwargs)
        self.bn = self._bn_cls(out_channels, momentum=bn_momentum) if bn else None
        self.relu = nn.ReLU(inplace=True) if relu else None
        self.init_weights()

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            if self.training and train_ops.bn_supported(x, self.bn):
                # batch-statistics BN + ReLU in two streaming passes of this repo's kernels (csrc/train_ops.cu) instead
                # of cuDNN's bn_fw_tr / bn_bw + separate ReLU passes; same values within fp32 rounding
                return train_ops.bn_relu_train(x, self.bn, self.relu is not None)
            x = self.bn(x)
        return x if self.relu is None else self.relu(x)

    def init_weights(self, init_fn=None):
        if init_fn is not None:
            init_fn(self.conv)
        if self.bn is not None:
            init_bn(self.bn)


class Conv1d(_ConvBnRelu):
    _conv_cls = nn.Conv1d
    _bn_cls = nn.BatchNorm1d


class Conv2d(_ConvBnRelu):
    _conv_cls = nn.Conv2d
    _bn_cls = nn.BatchNorm2d


class SharedMLP(nn.ModuleList):
    """Per-position MLP: a list of Conv1d/Conv2d(k=1) blocks; dropout only when training and dropout_prob > 0
    (F.dropout for ndim 1, F.dropout2d for ndim 2 -- nn/modules/mlp.py:95-106)."""

    def __init__(self, in_channels, mlp_channels, ndim=1, dropout_prob=0.0, bn=True, bn_momentum=0.1):
        super().__init__()
        if ndim not in (1, 2):
            raise ValueError('SharedMLP only supports ndim=(1, 2).')
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.ndim = ndim
        block = Conv1d if ndim == 1 else Conv2d
        c = in_channels
        for width in mlp_channels:
            self.append(block(c, width, 1, relu=True, bn=bn, bn_momentum=bn_momentum))
            c = width
        assert dropout_prob >= 0.0
        self.dropout_prob = dropout_prob

    def forward(self, x):
        if self.training and len(self) > 0 and conv_train.chain_supported(self, x):
            # train mode on CUDA: the whole MLP as one autograd function on this repo's tensor-core engine (conv_train.py)
            return conv_train.mlp_chain_train(self, x, pooled=False)
        drop = F.dropout if self.ndim == 1 else F.dropout2d
        for block in self:
            x = block(x)
            if self.training and self.dropout_prob > 0.0:
                x = drop(x, p=self.dropout_prob, training=True)
        return x

    def forward_max_over_neighbours(self, x):
        """torch.max(self(x), 3)[0] (modules.py:245).  In train mode on CUDA the last block's BatchNorm + ReLU and the
        max over the 64 neighbours are one pair of kernels: its (B, C, M, 64) activation is never written."""
        last = self[len(self) - 1]
        if (self.training and self.ndim == 2 and x.dim() == 4 and x.size(3) == 64 and self.dropout_prob == 0.0
                and conv_train.chain_supported(self, x)):
            return conv_train.mlp_chain_train(self, x, pooled=True)
        if (self.training and self.ndim == 2 and self.dropout_prob == 0.0 and last.bn is not None and x.is_cuda
                and os.environ.get("REGNET_TRAIN_UNFUSED_MAX", "0") != "1"):
            for i in range(len(self) - 1):
                x = self[i](x)
            x = last.conv(x)
            if train_ops.bn_relu_max64_supported(x, last.bn):
                return train_ops.bn_relu_max64_train(x, last.bn, last.relu is not None)
            x = last.bn(x)
            x = x if last.relu is None else last.relu(x)
            return train_ops.max_over_neighbours(x)
        return train_ops.max_over_neighbours(self.forward(x))

    def init_weights(self, init_fn=None):
        for block in self:
            block.init_weights(init_fn)

    def extra_repr(self):
        return 'dropout_prob={}'.format(self.dropout_prob) if self.dropout_prob > 0.0 else ''
