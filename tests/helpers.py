"""Shared parity definitions."""
import numpy as np
import torch

REL_TOL = 1e-4  # BASELINE.json north_star: MLP / pooled features within 1e-4 relative


def assert_features_close(got, want, tol=REL_TOL, what="feature"):
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    assert got.shape == want.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(want.shape)}"
    err = (got - want).abs()
    peak = want.abs().max().item()
    rms = want.pow(2).mean().sqrt().item()
    assert err.max().item() <= tol * max(peak, 1e-30), f"{what}: max err {err.max().item():.3e} vs {tol}*peak {peak:.3e}"
    bound = tol * want.abs() + tol * rms
    worst = (err - bound).max().item()
    assert worst <= 0, f"{what}: elementwise rtol/atol violated by {worst:.3e} (rms {rms:.3e})"


def robust_rel_err(got, want, outlier_frac):
    """max |got - want| / max |want| after discarding the `outlier_frac` largest errors.  For gradients that pass through
    ReLU / max decisions: an input within fp32 rounding of a decision boundary sends the gradient one way in fp32 and the
    other way in the float64 reference, which changes a handful of elements by O(1) without any arithmetic being wrong."""
    got = torch.as_tensor(got).double().cpu().flatten()
    want = torch.as_tensor(want).double().cpu().flatten()
    err = (got - want).abs()
    k = int(outlier_frac * err.numel())
    if k > 0:
        err = err.kthvalue(err.numel() - k)[0]
    else:
        err = err.max()
    return (err / want.abs().max().clamp_min(1e-30)).item()


def rel_l2(got, want):
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    return ((got - want).norm() / want.norm().clamp_min(1e-30)).item()


def rel_err(got, want):
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()


def bq_rowhash(bq, K):
    w = (np.arange(K, dtype=np.int64) * 2654435761 % 1000003 + 1)
    return (np.asarray(bq, dtype=np.int64) * w).sum(-1)


def region_net_fixture():
    import torch
    from regnet_for_3d_grasping_b200.gripper_region_network import GripperRegionNetwork
    from regnet_for_3d_grasping_b200.weights import seeded_state_like
    from conftest import golden
    ref = golden("ref_py_region_net.npz")
    net = GripperRegionNetwork(training=True, group_num=16, gripper_num=8, grasp_score_threshold=0.4, radius=0.06,
                               reg_channel=10).eval()
    sd = seeded_state_like(net.state_dict(), seed=int(ref["weight_seed"]))
    inp = {k: torch.from_numpy(ref[k]) for k in ("pc", "all_feature", "center_pc", "center_pc_index", "pc_group_index",
                                                 "pc_group", "pc_group_more_index", "pc_group_more")}
    return ref, net, sd, inp
