"""GPU parity, the shared-MLP contraction (conv1x1 + BN + ReLU [+ 64-neighbour max]) through the C ABI
(regnet_mlp_layer), both engines, against float64 torch on the same inputs.  Tolerance: 1e-4 relative
(BASELINE.json north_star), written in helpers.REL_TOL."""
import ctypes

import pytest
import torch

from helpers import assert_features_close, rel_err

pytestmark = pytest.mark.gpu


def _run(lib, X, W, scale, shift, pool, act, engine):
    P, cin = X.shape
    cout = W.shape[0]
    Y = torch.empty(P // pool if pool else P, cout, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.regnet_mlp_layer(p(X), p(W), p(scale), p(shift), P, cin, cout, pool, act, engine, p(Y), None)
    assert rc == 0, lib.regnet_last_error()
    return Y


def _ref(X, W, scale, shift, pool, act):
    y = X.double() @ W.double().t() * scale.double() + shift.double()
    if act == 1:
        y = y.relu()
    elif act == 2:
        y = y.sigmoid()
    if pool:
        y = y.view(-1, pool, y.shape[1]).max(1)[0]
    return y


SHAPES = [  # (P, cin, cout, pool)
    (64 * 50, 6, 128, 0), (64 * 50, 128, 128, 0), (64 * 50, 128, 256, 64),      # SA1-like
    (64 * 20, 259, 256, 0), (64 * 20, 256, 512, 64),                             # SA2-like
    (64 * 6, 515, 512, 0), (64 * 6, 512, 1024, 64),                              # SA3-like
    (1000, 1536, 1024, 0), (3001, 515, 256, 0), (777, 256, 128, 0),              # FP / seg, ragged P
    (130, 20, 12, 0), (64, 29, 24, 64), (1, 16, 8, 0),                           # tiny / odd channel counts
]


@pytest.mark.parametrize("engine", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("P,cin,cout,pool", SHAPES)
def test_mlp_layer_matches_float64(lib_path, engine, P, cin, cout, pool):
    from regnet_for_3d_grasping_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(P + cin + cout)
    X = (torch.randn(P, cin, generator=g) * torch.rand(1, cin, generator=g) * 3).cuda()
    W = ((torch.rand(cout, cin, generator=g) * 2 - 1) / cin ** 0.5).cuda()
    scale = (torch.rand(cout, generator=g) + 0.5).cuda() * torch.where(torch.rand(cout, generator=g) < 0.2, -1.0, 1.0).cuda()
    shift = (torch.randn(cout, generator=g) * 0.1).cuda()
    Y = _run(lib, X, W, scale, shift, pool, 1, engine)
    assert_features_close(Y, _ref(X, W, scale, shift, pool, 1), what=f"engine {engine} {P}x{cin}->{cout} pool {pool}")


@pytest.mark.parametrize("engine", [1, 0], ids=["simt", "tcgen05"])
def test_mlp_layer_activations_and_no_bn(lib_path, engine):
    from regnet_for_3d_grasping_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    X = torch.randn(500, 128, generator=g).cuda()
    W = (torch.randn(64, 128, generator=g) / 11).cuda()
    one, zero = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    for act in (0, 1, 2):
        Y = _run(lib, X, W, one, zero, 0, act, engine)
        assert_features_close(Y, _ref(X, W, one, zero, 0, act), what=f"act {act}")


def test_split_bf16_error_budget(lib_path):
    """Report (and bound) the error of the 3-product bf16 split on the deepest contraction of the net."""
    from regnet_for_3d_grasping_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(11)
    X = torch.randn(2048, 1536, generator=g).cuda()
    W = ((torch.rand(1024, 1536, generator=g) * 2 - 1) / 1536 ** 0.5).cuda()
    one, zero = torch.ones(1024, device="cuda"), torch.zeros(1024, device="cuda")
    ref = _ref(X, W, one, zero, 0, 0)
    e_tc = rel_err(_run(lib, X, W, one, zero, 0, 0, 0), ref)
    e_simt = rel_err(_run(lib, X, W, one, zero, 0, 0, 1), ref)
    print(f"\nmax-rel error K=1536: tcgen05 split-bf16 {e_tc:.3e}, fp32 SIMT {e_simt:.3e}")
    assert e_tc < 2e-5 and e_simt < 2e-5
