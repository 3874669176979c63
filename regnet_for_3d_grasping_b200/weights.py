"""Seeded ScoreNet state dicts (names/shapes of SURVEY.md Appendix B) for tests and benchmarks.

The reference ships no weights (test.py:33-34 points at unreleased files), so parity tests and bench.py use
random weights with the reference's default init distributions and, by default, NON-trivial BN statistics so that
BatchNorm bugs cannot hide behind the identity."""
import torch

SA_CHANNELS = ((128, 128, 256), (256, 256, 512), (512, 512, 1024))   # multi_model/utils/pointnet2.py:43
FP_CHANNELS = ((1024, 1024), (512, 512), (256, 256, 256))            # :44
SEG_CHANNELS = (512, 256, 256, 128)                                  # :46


def scorenet_state_shapes(input_chann=6):
    """name -> shape of every tensor in ScoreNetwork.state_dict(); 127 tensors."""
    shapes = {}

    def layer(prefix, cin, cout, ndim):
        shapes[prefix + ".conv.weight"] = (cout, cin) + (1,) * ndim
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + ".bn." + k] = (cout,)
        shapes[prefix + ".bn.num_batches_tracked"] = ()

    root = "extrat_featurePN2."
    feat = input_chann - 3
    inter = [feat]
    for i, chans in enumerate(SA_CHANNELS):
        cin = feat + 3
        for j, cout in enumerate(chans):
            layer(f"{root}sa_modules.{i}.mlp.{j}", cin, cout, 2)
            cin = cout
        feat = chans[-1]
        inter.append(feat)
    for i, chans in enumerate(FP_CHANNELS):
        cin = feat + inter[-2 - i]
        for j, cout in enumerate(chans):
            layer(f"{root}fp_modules.{i}.mlp.{j}", cin, cout, 1)
            cin = cout
        feat = chans[-1]
    cin = feat
    for j, cout in enumerate(SEG_CHANNELS):
        layer(f"{root}mlp.{j}", cin, cout, 1)
        cin = cout
    shapes[root + "conv_score.weight"] = (1, cin, 1)
    shapes[root + "conv_score.bias"] = (1,)
    for k in ("weight", "bias", "running_mean", "running_var"):
        shapes[root + "bn_score." + k] = (1,)
    shapes[root + "bn_score.num_batches_tracked"] = ()
    return shapes


def random_scorenet_state(seed=0, randomize_bn=True, dtype=torch.float32):
    """Seeded weights with the reference's default init distributions (kaiming-uniform conv as nn.Conv*d,
    BN gamma=1/beta=0 per nn/init.py:4-8) and, by default, NON-trivial BN statistics/affine so that BN
    bugs cannot hide behind the identity."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape in scorenet_state_shapes().items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(0, dtype=torch.int64)
        elif name.endswith("conv.weight") or name.endswith("conv_score.weight"):
            fan_in = shape[1]
            bound = 1.0 / fan_in ** 0.5
            sd[name] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul(bound).to(dtype)
        elif name.endswith("conv_score.bias"):
            sd[name] = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1).mul(1.0 / 128 ** 0.5).to(dtype)
        elif name.endswith("running_var"):
            v = torch.rand(shape, generator=g, dtype=torch.float64) * 1.5 + 0.5 if randomize_bn else torch.ones(shape, dtype=torch.float64)
            sd[name] = v.to(dtype)
        elif name.endswith("running_mean"):
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1 if randomize_bn else torch.zeros(shape, dtype=torch.float64)
            sd[name] = v.to(dtype)
        elif name.endswith("bn.weight") or name.endswith("bn_score.weight"):
            v = torch.rand(shape, generator=g, dtype=torch.float64) * 1.0 + 0.5 if randomize_bn else torch.ones(shape, dtype=torch.float64)
            sd[name] = v.to(dtype)
        elif name.endswith("bn.bias") or name.endswith("bn_score.bias"):
            v = torch.randn(shape, generator=g, dtype=torch.float64) * 0.1 if randomize_bn else torch.zeros(shape, dtype=torch.float64)
            sd[name] = v.to(dtype)
        else:  # pragma: no cover
            raise KeyError(name)
    return sd


def seeded_state_like(state, seed=0):
    """Deterministic values for every tensor of a state dict, derived from (seed, key name) only -- independent of module
    registration order, so a fixture generator and a test can agree on weights without storing them.
    Conv/linear weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); biases small; BN gamma in [0.5,1.5], var in [0.5,1.5]."""
    import zlib
    out = {}
    for name, ref in state.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7fffffff)
        shape = tuple(ref.shape)
        if name.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=ref.dtype)
        elif name.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif name.endswith(".weight") and len(shape) == 1:
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith(".weight"):
            v = (torch.rand(shape, generator=g) * 2 - 1) / shape[1] ** 0.5
        else:
            v = torch.randn(shape, generator=g) * 0.05
        out[name] = v.to(ref.dtype)
    return out


def seeded_region_state(state, seed=0, residual_scale=0.1):
    """seeded_state_like for a GripperRegionNetwork whose first-stage regression head is scaled down: an untrained head
    emits O(1) residuals, i.e. grasp centres 6 cm (one gripper depth) away from their anchor, and the closing boxes of such
    proposals are empty on a real cloud; with residuals of O(0.1) the proposals stay on the surface, like a trained
    network's, and the refine stage has something to refine.  Used by the virtual-data fixture and its GPU test."""
    out = seeded_state_like(state, seed)
    # 40 regression channels = 4 anchors x (3 centre, 3 axis, 1 angle residuals + 3 scores): scale the 7 residuals only,
    # the predicted scores keep their spread (they are thresholded in the first-stage selection)
    scale = torch.ones(4, 10)
    scale[:, :7] = residual_scale
    for name in ("extrat_feature_region.bn_reg4.weight", "extrat_feature_region.bn_reg4.bias"):
        out[name] = out[name] * scale.view(-1)
    return out
