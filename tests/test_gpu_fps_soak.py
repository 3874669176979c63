"""Randomised soak of furthest-point sampling against the C oracle: sizes across every launch-shape bracket of the policy
(one CTA, 4-CTA and 8-CTA clusters, multi-pick rounds), ragged point counts, duplicated points (zero distances: the
"repeat the previous pick" rule), quantised coordinates (heavy ties) -- indices must be identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(rng, B, N, kind):
    if kind == "uniform":
        pts = rng.random((B, N, 3), dtype=np.float32)
    elif kind == "quantised":                      # coordinates on a coarse lattice: many exactly equal distances
        pts = (rng.integers(0, 12, (B, N, 3)) / 8.0).astype(np.float32)
    elif kind == "duplicates":                     # a third of the points are copies of other points
        pts = rng.random((B, N, 3), dtype=np.float32)
        src = rng.integers(0, N, (B, N // 3))
        dst = rng.integers(0, N, (B, N // 3))
        for b in range(B):
            pts[b, dst[b]] = pts[b, src[b]]
    else:                                          # thin slab (table-top like)
        pts = rng.random((B, N, 3), dtype=np.float32) * np.array([1.0, 0.8, 0.02], dtype=np.float32)
    return pts


@pytest.mark.parametrize("seed", range(6))
def test_fps_random_shapes_match_the_oracle(lib_path, oracle, seed):
    from regnet_for_3d_grasping_b200 import pn2_ext
    rng = np.random.default_rng(1000 + seed)
    brackets = [(17, 512), (513, 2048), (2049, 12288), (12289, 30000)]
    for lo, hi in brackets:
        N = int(rng.integers(lo, hi + 1))
        B = int(rng.integers(1, 4))
        M = int(rng.integers(1, max(2, min(N, 1500))))
        kind = ["uniform", "quantised", "duplicates", "slab"][int(rng.integers(0, 4))]
        pts = _cloud(rng, B, N, kind)
        xyz = torch.from_numpy(pts).permute(0, 2, 1).contiguous()
        want = oracle.farthest_point_sample(xyz, M)
        got = pn2_ext.farthest_point_sample(xyz.cuda(), M).cpu()
        assert torch.equal(got, want), f"N={N} B={B} M={M} {kind}: first difference at {int((got != want).nonzero()[0][1])}"
