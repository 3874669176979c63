"""Region / refine heads of REGNet: mirrors of
multi_model/utils/pointnet2.py:123-197 (PointNet2TwoStage) and :199-254 (PointNet2Refine) with the same constructor
arguments, attribute names (=> state-dict keys of Appendix B) and outputs.

The reference feeds these heads a materialised (M, 256, G) gather of per-point features and max-pools it inside the
module (MaxPool1d).  `forward` keeps that contract; `forward_pooled` takes the already pooled (M, 256) features that
regnet_for_3d_grasping_b200.region.gather_max produces in one pass over all_feature (no (M, G, 256) tensor at all).
In eval mode on CUDA the per-centre MLPs run on this repo's tcgen05 engine (regnet_linear_planes: split-bf16 GEMM with the
BatchNorm folded into the epilogue, activations handed from layer to layer as bf16 planes): 7 launches for the region head
and 5 for the refine head instead of ~20 / ~14 cuBLAS + BatchNorm + ReLU launches.  In train mode (BN batch statistics,
autograd) they stay torch modules -- a few thousand rows of tiny 1x1 convolutions.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _conv1x1(conv, x2d):
    """A Conv1d(kernel 1) applied to (M, C_in) rows as an fp32 GEMM.  cuDNN convolutions default to TF32 on this
    hardware (torch.backends.cudnn.allow_tf32), which costs ~1e-3 of relative accuracy; torch.matmul stays fp32."""
    return torch.addmm(conv.bias, x2d, conv.weight.view(conv.out_channels, conv.in_channels).t())


def _p(t):
    return None if t is None else t.data_ptr()


class _FoldedLayer:
    """conv (1x1, with bias) + BatchNorm (eval) as planes of W and a per-channel (scale, shift)."""

    def __init__(self, conv, bn):
        from .conv_train import split_weight
        w = conv.weight.detach().float().reshape(conv.out_channels, conv.in_channels)
        g, b = bn.weight.detach().float(), bn.bias.detach().float()
        scale = g / torch.sqrt(bn.running_var.detach().float() + bn.eps)
        self.scale = scale.contiguous()
        self.shift = (b + (conv.bias.detach().float() - bn.running_mean.detach().float()) * scale).contiguous()
        self.w_hi, self.w_lo = split_weight(w)
        self.cin, self.cout = conv.in_channels, conv.out_channels


class _NativeHeads:
    """Eval-mode executor shared by the two heads: folded layers cached per parameter version."""

    def __init__(self, module, pairs):
        self.module, self.pairs, self.key, self.layers = module, pairs, None, None

    def usable(self, x):
        return (not self.module.training and x.is_cuda and x.dtype == torch.float32 and x.shape[0] > 0
                and os.environ.get("REGNET_HEADS_TORCH", "0") != "1")

    def _layers(self):
        tensors = []
        for conv, bn in self.pairs:
            c, b = getattr(self.module, conv), getattr(self.module, bn)
            tensors += [c.weight, c.bias, b.weight, b.bias, b.running_mean, b.running_var]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key != self.key:
            with torch.no_grad():
                self.layers = {conv: _FoldedLayer(getattr(self.module, conv), getattr(self.module, bn)) for conv, bn in self.pairs}
            self.key = key
        return self.layers

    @staticmethod
    def run(layer, x_hi, x_lo, rows, act, planes):
        """One layer: planes in -> planes (planes=True) or fp32 (rows, round_up(cout, 4)) out."""
        lib = _lib.load()
        dev = x_hi.device
        if planes:
            ld = (layer.cout + 7) // 8 * 8
            o_hi = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
            o_lo = torch.empty(rows, ld, dtype=torch.bfloat16, device=dev)
            _lib.check(lib.regnet_linear_planes(_p(x_hi), _p(x_lo), x_hi.shape[1], rows, layer.cin, _p(layer.w_hi), _p(layer.w_lo),
                                                layer.w_hi.shape[1], layer.cout, _p(layer.scale), _p(layer.shift), act, None, 0,
                                                _p(o_hi), _p(o_lo), ld, _lib.current_stream_ptr()))
            return o_hi, o_lo
        ld = (layer.cout + 3) // 4 * 4
        out = torch.empty(rows, ld, dtype=torch.float32, device=dev)
        _lib.check(lib.regnet_linear_planes(_p(x_hi), _p(x_lo), x_hi.shape[1], rows, layer.cin, _p(layer.w_hi), _p(layer.w_lo),
                                            layer.w_hi.shape[1], layer.cout, _p(layer.scale), _p(layer.shift), act, _p(out), ld,
                                            None, None, 0, _lib.current_stream_ptr()))
        return out[:, :layer.cout]


class PointNet2TwoStage(nn.Module):
    def __init__(self, num_points, input_chann, k_cls, k_reg, k_reg_theta, add_channel_flag=False):
        super().__init__()
        self.num_points = num_points
        self.k_reg = k_reg
        self.k_cls = k_cls
        self.k_reg_no_anchor = self.k_reg // self.k_cls
        self.k_reg_theta = k_reg_theta
        self.conv = nn.Conv1d(256 * (3 if add_channel_flag else 1), 1024, 1)
        self.bn = nn.BatchNorm1d(1024)
        self.conv_cls2 = nn.Conv1d(1024, 256, 1)
        self.conv_cls3 = nn.Conv1d(256, 128, 1)
        self.linear_cls = nn.Linear(128, self.k_cls)          # present (and unused) in the reference: keeps the keys
        self.conv_cls4 = nn.Conv1d(128, self.k_cls, 1)
        self.bn_cls2 = nn.BatchNorm1d(256)
        self.bn_cls3 = nn.BatchNorm1d(128)
        self.bn_cls4 = nn.BatchNorm1d(self.k_cls)
        self.conv_reg2 = nn.Conv1d(1024, 256, 1)
        self.conv_reg3 = nn.Conv1d(256, 128, 1)
        self.conv_reg4 = nn.Conv1d(128, self.k_reg, 1)
        self.bn_reg2 = nn.BatchNorm1d(256)
        self.bn_reg3 = nn.BatchNorm1d(128)
        self.bn_reg4 = nn.BatchNorm1d(self.k_reg)
        self.mp1 = nn.MaxPool1d(num_points)
        self.ap = nn.AdaptiveAvgPool1d(1)
        self.sigmod = nn.Sigmoid()

    _PAIRS = (("conv", "bn"), ("conv_cls2", "bn_cls2"), ("conv_cls3", "bn_cls3"), ("conv_cls4", "bn_cls4"),
              ("conv_reg2", "bn_reg2"), ("conv_reg3", "bn_reg3"), ("conv_reg4", "bn_reg4"))

    def _heads_native(self, m):
        from .conv_train import split_planes
        if getattr(self, "_native", None) is None or self._native.module is not self:
            object.__setattr__(self, "_native", _NativeHeads(self, self._PAIRS))
        L = self._native._layers()
        rows = m.shape[0]
        run = _NativeHeads.run
        with torch.cuda.device(m.device):
            hi, lo = split_planes(m.contiguous())
            x = run(L["conv"], hi, lo, rows, 1, True)
            c = run(L["conv_cls2"], x[0], x[1], rows, 1, True)
            c = run(L["conv_cls3"], c[0], c[1], rows, 1, True)
            x_cls = run(L["conv_cls4"], c[0], c[1], rows, 0, False)
            r = run(L["conv_reg2"], x[0], x[1], rows, 1, True)
            r = run(L["conv_reg3"], r[0], r[1], rows, 1, True)
            r = run(L["conv_reg4"], r[0], r[1], rows, 0, False)
        x_reg = r.reshape(rows, -1, self.k_reg_no_anchor)
        x_reg[:, :, 7:] = self.sigmod(x_reg[:, :, 7:])
        return x_cls.contiguous(), x_reg

    def _heads(self, mp_x):
        m = mp_x.reshape(mp_x.shape[0], -1)                        # (M, C, 1) -> (M, C): every layer is a 1x1 convolution
        if getattr(self, "_native", None) is None:
            object.__setattr__(self, "_native", _NativeHeads(self, self._PAIRS))
        if self._native.usable(m):
            return self._heads_native(m)
        x = F.relu(self.bn(_conv1x1(self.conv, m)))
        c = F.relu(self.bn_cls2(_conv1x1(self.conv_cls2, x)))
        c = F.relu(self.bn_cls3(_conv1x1(self.conv_cls3, c)))
        x_cls = self.bn_cls4(_conv1x1(self.conv_cls4, c))
        r = F.relu(self.bn_reg2(_conv1x1(self.conv_reg2, x)))
        r = F.relu(self.bn_reg3(_conv1x1(self.conv_reg3, r)))
        r = self.bn_reg4(_conv1x1(self.conv_reg4, r))
        x_reg = r.view(r.size(0), -1, self.k_reg_no_anchor)
        x_reg[:, :, 7:] = self.sigmod(x_reg[:, :, 7:])      # scores in (0,1); in place like the reference (:189)
        return x_cls, x_reg

    def forward(self, xyz, feature):
        """xyz (M, 256, G) gathered per-point features [, feature (M, C')] -> x_cls (M,k_cls), x_reg (M,k_cls,k_reg/k_cls),
        mp_x (M, 256[+C'], 1)."""
        mp_x = self.mp1(xyz)
        if feature is not None:
            mp_x = torch.cat((mp_x, feature.view(feature.shape[0], feature.shape[1], 1)), dim=1)
        x_cls, x_reg = self._heads(mp_x)
        return x_cls, x_reg, mp_x

    def __getstate__(self):           # the executor caches device tensors and points back at the module: never pickled
        state = self.__dict__.copy()
        state.pop("_native", None)
        return state

    def forward_pooled(self, pooled):
        """pooled (M, 256) = max over each centre's group (region.gather_max) -> same outputs as forward()."""
        mp_x = pooled.unsqueeze(-1)
        x_cls, x_reg = self._heads(mp_x)
        return x_cls, x_reg, mp_x


class PointNet2Refine(nn.Module):
    def __init__(self, num_points=2500, input_chann=3, k_cls=2, k_reg=8):
        super().__init__()
        self.num_points = num_points
        self.k_reg = k_reg
        self.k_cls = k_cls
        self.conv_formal = nn.Conv1d(384, 1024, 1)
        self.bn_formal = nn.BatchNorm1d(1024)
        self.conv_formal_cls2 = nn.Conv1d(1024, 128, 1)
        self.conv_formal_cls3 = nn.Conv1d(128, self.k_cls, 1)
        self.bn_formal_cls2 = nn.BatchNorm1d(128)
        self.bn_formal_cls3 = nn.BatchNorm1d(self.k_cls)
        self.conv_formal_reg2 = nn.Conv1d(1024, 128, 1)
        self.conv_formal_reg3 = nn.Conv1d(128, self.k_reg, 1)
        self.bn_formal_reg2 = nn.BatchNorm1d(128)
        self.bn_formal_reg3 = nn.BatchNorm1d(self.k_reg)
        self.mp1 = nn.MaxPool1d(num_points)
        self.ap = nn.AdaptiveAvgPool1d(1)
        self.sigmoid = nn.Sigmoid()

    _PAIRS = (("conv_formal", "bn_formal"), ("conv_formal_cls2", "bn_formal_cls2"), ("conv_formal_cls3", "bn_formal_cls3"),
              ("conv_formal_reg2", "bn_formal_reg2"), ("conv_formal_reg3", "bn_formal_reg3"))

    def _heads_native(self, m):
        from .conv_train import split_planes
        L = self._native._layers()
        rows = m.shape[0]
        run = _NativeHeads.run
        with torch.cuda.device(m.device):
            hi, lo = split_planes(m.contiguous())
            x = run(L["conv_formal"], hi, lo, rows, 1, True)
            c = run(L["conv_formal_cls2"], x[0], x[1], rows, 1, True)
            c = run(L["conv_formal_cls3"], c[0], c[1], rows, 0, False)
            r = run(L["conv_formal_reg2"], x[0], x[1], rows, 1, True)
            r = run(L["conv_formal_reg3"], r[0], r[1], rows, 0, False)
        return c.contiguous(), r.contiguous()

    def _heads(self, x):
        m = x.reshape(x.shape[0], -1)
        if getattr(self, "_native", None) is None:
            object.__setattr__(self, "_native", _NativeHeads(self, self._PAIRS))
        if self._native.usable(m):
            return self._heads_native(m)
        x = F.relu(self.bn_formal(_conv1x1(self.conv_formal, x.reshape(x.shape[0], -1))))
        c = F.relu(self.bn_formal_cls2(_conv1x1(self.conv_formal_cls2, x)))
        c = self.bn_formal_cls3(_conv1x1(self.conv_formal_cls3, c))
        r = F.relu(self.bn_formal_reg2(_conv1x1(self.conv_formal_reg2, x)))
        r = self.bn_formal_reg3(_conv1x1(self.conv_formal_reg3, r))
        return c, r

    def forward(self, gripper_feature, group_feature):
        """gripper_feature (M', 256, 64) [, group_feature (M', 128)] -> x_cls (M',k_cls), x_reg (M',k_reg)."""
        x = self.mp1(gripper_feature)
        if group_feature is not None:
            x = torch.cat((x, group_feature.view(group_feature.shape[0], group_feature.shape[1], 1)), dim=1)
        return self._heads(x)

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_native", None)
        return state

    def forward_pooled(self, pooled, group_feature):
        """pooled (M', 256) from region.gather_max over the 64 closing-box points; group_feature (M', 128)."""
        x = pooled.unsqueeze(-1)
        if group_feature is not None:
            x = torch.cat((x, group_feature.view(group_feature.shape[0], group_feature.shape[1], 1)), dim=1)
        return self._heads(x)
