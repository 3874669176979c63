"""Python handle on the native ScoreNet plan (include/regnet_b200.h section 2, csrc/scorenet.cu).

`ScoreNetPlan` owns a `regnet_scorenet*`; `bind_state()` folds the reference's conv/BN parameters
(SURVEY.md Appendix B names) into per-layer (W, scale, shift) on the device; `forward()` runs the whole
PointNet2Seg forward (multi_model/utils/pointnet2.py:86-121) on torch's current stream.
"""
import ctypes

import torch

from . import _lib

# reference constants, multi_model/utils/pointnet2.py:40-46
NUM_CENTROIDS = (5120, 1024, 256)
RADIUS = (0.02, 0.08, 0.32)
NUM_NEIGHBOURS = (64, 64, 64)
BN_EPS = 1e-5

_STAGE_PREFIX = [("sa_modules.0.mlp", 3), ("sa_modules.1.mlp", 3), ("sa_modules.2.mlp", 3),
                 ("fp_modules.0.mlp", 2), ("fp_modules.1.mlp", 2), ("fp_modules.2.mlp", 3), ("mlp", 4)]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def fold_bn(sd, prefix, eps=BN_EPS):
    """(W (cout,cin), scale, shift) of one `<prefix>.conv` + `<prefix>.bn` block in eval mode:
    y = (Wx - mean) / sqrt(var + eps) * gamma + beta = scale * (Wx) + shift."""
    w = sd[prefix + ".conv.weight"].detach().float()
    w = w.reshape(w.shape[0], w.shape[1]).contiguous()
    g, b = sd[prefix + ".bn.weight"].detach().float(), sd[prefix + ".bn.bias"].detach().float()
    m, v = sd[prefix + ".bn.running_mean"].detach().float(), sd[prefix + ".bn.running_var"].detach().float()
    scale = g / torch.sqrt(v + eps)
    shift = b - m * scale
    return w, scale.contiguous(), shift.contiguous()


class ScoreNetPlan:
    def __init__(self, batch, num_points, device, engine=None, num_centroids=NUM_CENTROIDS, radius=RADIUS,
                 side_stream=3):
        lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("ScoreNetPlan needs a CUDA device; there is no CPU path")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if engine is None:
            engine = _lib.ENGINE_TC
        self.engine = engine
        self.batch, self.num_points = int(batch), int(num_points)
        self.num_centroids = tuple(int(x) for x in num_centroids)
        cfg = _lib.ScoreNetConfig()
        cfg.batch, cfg.num_points = self.batch, self.num_points
        for i in range(3):
            cfg.num_centroids[i] = self.num_centroids[i]
            cfg.radius[i] = float(radius[i])
            cfg.num_neighbours[i] = NUM_NEIGHBOURS[i]
        cfg.engine = engine
        # 0: single stream; 1: whole geometry chain on the side stream; 2: only FPS on side streams;
        # 3 (default): FPS on side streams + ball query of levels 1-2 and 3-NN behind their FPS (regnet_b200.h)
        cfg.use_side_stream = int(side_stream)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.regnet_scorenet_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self._bound_key = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().regnet_scorenet_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def workspace_bytes(self):
        return int(_lib.load().regnet_scorenet_workspace_bytes(self._h))

    @property
    def launch_count(self):
        return int(_lib.load().regnet_scorenet_launch_count(self._h))

    def bind_state(self, sd, root="extrat_featurePN2.", key=None):
        """Upload folded weights.  `sd` maps the reference's state-dict names to tensors (any device)."""
        if key is not None and key == self._bound_key:
            return
        lib = _lib.load()
        stream = _lib.current_stream_ptr()
        keep = []
        with torch.cuda.device(self.device):
            for stage, (prefix, nl) in enumerate(_STAGE_PREFIX):
                for layer in range(nl):
                    w, scale, shift = fold_bn(sd, f"{root}{prefix}.{layer}")
                    w, scale, shift = (t.to(self.device) for t in (w, scale, shift))
                    keep += [w, scale, shift]
                    _lib.check(lib.regnet_scorenet_set_layer(self._h, stage, layer, w.shape[1], w.shape[0], _p(w),
                                                             _p(scale), _p(shift), stream))
            # score head: Conv1d(128,1) WITH bias -> BatchNorm1d(1) -> sigmoid (pointnet2.py:82-84,117-119)
            w = sd[root + "conv_score.weight"].detach().float().reshape(1, -1).contiguous().to(self.device)
            bias = sd[root + "conv_score.bias"].detach().float().to(self.device)
            g, b = sd[root + "bn_score.weight"].detach().float().to(self.device), sd[root + "bn_score.bias"].detach().float().to(self.device)
            m, v = sd[root + "bn_score.running_mean"].detach().float().to(self.device), sd[root + "bn_score.running_var"].detach().float().to(self.device)
            scale = (g / torch.sqrt(v + BN_EPS)).contiguous()
            shift = (b + (bias - m) * scale).contiguous()
            keep += [w, scale, shift]
            _lib.check(lib.regnet_scorenet_set_layer(self._h, 7, 0, 128, 1, _p(w), _p(scale), _p(shift), stream))
            torch.cuda.current_stream().synchronize()  # `keep` may be freed after this
        self._bound_key = key

    def forward(self, pc, all_feature=None, score=None):
        """pc (B,N,>=6) fp32 CUDA -> (all_feature (B,N,256), score (B,N)).  Asynchronous on the current stream."""
        if pc.device != self.device or pc.dtype != torch.float32:
            raise RuntimeError("pc must be a float32 tensor on the plan's device")
        if pc.dim() != 3 or pc.size(0) != self.batch or pc.size(1) != self.num_points or pc.size(2) < 6:
            raise RuntimeError(f"pc must be ({self.batch}, {self.num_points}, >=6), got {tuple(pc.shape)}")
        if pc.size(2) != 6 or not pc.is_contiguous():
            pc = pc[:, :, :6].contiguous()
        if all_feature is None:
            all_feature = torch.empty(self.batch, self.num_points, 256, dtype=torch.float32, device=self.device)
        if score is None:
            score = torch.empty(self.batch, self.num_points, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().regnet_scorenet_forward(self._h, _p(pc), _p(all_feature), _p(score),
                                                           _lib.current_stream_ptr()))
        self._last_pc = pc  # keep the (possibly copied) input alive until the stream has consumed it
        return all_feature, score

    def prefetch(self, pc):
        """Enqueue the geometry chain (FPS / ball query / 3-NN) of a FUTURE forward(pc) on the plan's side stream so
        that it overlaps the MLPs of the forward issued next.  `pc` must be the very tensor (contiguous (B,N,6) fp32)
        later passed to forward() and must not change in between."""
        if pc.device != self.device or pc.dtype != torch.float32 or not pc.is_contiguous() or tuple(pc.shape) != (self.batch, self.num_points, 6):
            raise RuntimeError("prefetch needs the contiguous (B, N, 6) float32 tensor that forward() will receive")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().regnet_scorenet_prefetch(self._h, _p(pc), _lib.current_stream_ptr()))
        # The library keeps only the pointer, and a parked (deferred) prefetch reads it later: hold the tensors of the (at
        # most two) outstanding prefetches so that a caller who drops its reference cannot leave that pointer dangling.
        held = getattr(self, "_prefetched", [])
        self._prefetched = (held + [pc])[-2:]

    def set_option(self, name, value):
        _lib.check(_lib.load().regnet_scorenet_set_option(self._h, name.encode(), int(value)))

    def geometry(self, pc):
        """FPS / ball query / 3-NN of every level for `pc` (B,N,6) -- the search results the training path needs, computed
        by the plan's kernels on its side streams (or taken from a matching prefetch()).  Returns a dict of fresh tensors:
        new_xyz[i] (B,3,M_i), nbr[i] (B,M_i,64) int64, nn_idx[f] / nn_w[f] (B,Nd_f,3) for the three FP modules."""
        if pc.device != self.device or pc.dtype != torch.float32 or not pc.is_contiguous() or tuple(pc.shape) != (self.batch, self.num_points, 6):
            raise RuntimeError("geometry needs a contiguous (B, N, 6) float32 tensor on the plan's device")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().regnet_scorenet_geometry(self._h, _p(pc), _lib.current_stream_ptr()))
        B, N, M = self.batch, self.num_points, self.num_centroids
        nd = (M[1], M[0], N)
        out = {"new_xyz": [], "nbr": [], "nn_idx": [], "nn_w": []}
        for i in range(3):
            out["new_xyz"].append(self.intermediate(f"xyz{i}", torch.float32, (B, 3, M[i])))
            out["nbr"].append(self.intermediate(f"bq{i}", torch.int32, (B, M[i], 64)).long())
            out["nn_idx"].append(self.intermediate(f"nn{i}", torch.int32, (B, nd[i], 3)).long())
            out["nn_w"].append(self.intermediate(f"nnw{i}", torch.float32, (B, nd[i], 3)))
        return out

    def join_prefetch(self):
        """Make the current stream wait for every outstanding prefetch()."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().regnet_scorenet_join_prefetch(self._h, _lib.current_stream_ptr()))

    def profile_forward(self, pc):
        """One forward with every launch bracketed by CUDA events (serial, caller's stream).
        Returns [(label, milliseconds), ...] in launch order."""
        lib = _lib.load()
        _lib.check(lib.regnet_scorenet_set_profiling(self._h, 1))
        try:
            self.forward(pc)
            buf = ctypes.create_string_buffer(1 << 16)
            _lib.check(lib.regnet_scorenet_profile(self._h, buf, len(buf)))
        finally:
            _lib.check(lib.regnet_scorenet_set_profiling(self._h, 0))
        return [(label, ms) for label, ms, _ in self._parse_profile(buf)]

    @staticmethod
    def _parse_profile(buf):
        out = []
        for line in buf.value.decode().splitlines():
            label, ms, t0 = line.rsplit(" ", 2)
            out.append((label, float(ms), float(t0)))
        return out

    def timeline(self, fn):
        """Run fn() (any sequence of prefetch()/forward() calls) with an event pair around every launch but WITHOUT
        serialising the streams; returns [(label, duration_ms, start_ms_since_first_launch), ...]."""
        lib = _lib.load()
        _lib.check(lib.regnet_scorenet_set_profiling(self._h, 2))
        try:
            fn()
            buf = ctypes.create_string_buffer(1 << 20)
            _lib.check(lib.regnet_scorenet_profile(self._h, buf, len(buf)))
        finally:
            _lib.check(lib.regnet_scorenet_set_profiling(self._h, 0))
        return self._parse_profile(buf)

    def intermediate(self, name, dtype, shape):
        """Copy of an intermediate of the last forward (see regnet_scorenet_intermediate)."""
        ptr, numel = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(_lib.load().regnet_scorenet_intermediate(self._h, name.encode(), ctypes.byref(ptr), ctypes.byref(numel)))
        src = _UnownedTensor.wrap(ptr.value, numel.value, dtype, self.device)
        return src.clone().view(*shape)


class _UnownedTensor:
    """Zero-copy torch view over device memory owned by the native plan (via __cuda_array_interface__)."""

    def __init__(self, ptr, numel, dtype):
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}

    @staticmethod
    def wrap(ptr, numel, dtype, device):
        with torch.cuda.device(device):
            return torch.as_tensor(_UnownedTensor(ptr, numel, dtype), device=device)
