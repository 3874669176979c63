"""Seeded synthetic point clouds shaped like REGNet's inputs (numpy only, no torch).

The reference trains/tests on 25 600-point RGB table-top scenes (train.py:70 `all_points_num`,
dataset_utils/scoredataset.py:60-81); the dataset itself is not distributed, so benchmarks and parity tests
use these generators:

  table_scene   60 % of the points on a 0.8 x 0.7 m table plane at z = 0.75 (N(0, 1 mm) depth noise), 40 % on
                the camera-facing surfaces of 8-12 boxes / spheres / cylinders (3-12 cm) standing on it;
                colours per object + N(0, .02); optional fraction of exact duplicates (the dataset resamples
                with replacement when a scene has < 25 600 points, scoredataset.py:71-72) to exercise ties.
  uniform_cube  U[0,1]^3, U[0,1] colours.
  lattice       points on a coarse integer lattice scaled to metres: many exact distance ties (adversarial
                for FPS / ball-query / 3-NN tie-breaking).
"""
import numpy as np


def table_scene(seed, n=25600, dup_frac=0.01):
    rng = np.random.default_rng(seed)
    n_plane = int(round(n * 0.6))
    n_obj = n - n_plane
    pts = np.empty((n, 3), dtype=np.float64)
    col = np.empty((n, 3), dtype=np.float64)
    # table plane
    pts[:n_plane, 0] = rng.uniform(-0.4, 0.4, n_plane)
    pts[:n_plane, 1] = rng.uniform(-0.35, 0.35, n_plane)
    pts[:n_plane, 2] = 0.75 + rng.normal(0.0, 0.001, n_plane)
    col[:n_plane] = np.array([0.55, 0.45, 0.35]) + rng.normal(0, 0.02, (n_plane, 3))
    # objects: camera looks down the -z axis from above with a small tilt, so we keep top + two side faces
    k = int(rng.integers(8, 13))
    share = rng.dirichlet(np.ones(k) * 4.0)
    counts = np.floor(share * n_obj).astype(int)
    counts[0] += n_obj - counts.sum()
    o = n_plane
    for i in range(k):
        m = int(counts[i])
        if m == 0:
            continue
        cx, cy = rng.uniform(-0.3, 0.3), rng.uniform(-0.25, 0.25)
        sx, sy, sz = rng.uniform(0.03, 0.12, 3)
        kind = int(rng.integers(0, 3))
        base = rng.uniform(0.1, 0.9, 3)
        u = rng.uniform(0, 1, m)
        v = rng.uniform(0, 1, m)
        face = rng.integers(0, 3, m)
        p = np.empty((m, 3))
        if kind == 0:      # box: top face / +x face / -y face
            top = face == 0
            fx = face == 1
            fy = face == 2
            p[top] = np.stack([cx + (u[top] - .5) * sx, cy + (v[top] - .5) * sy, np.full(top.sum(), 0.75 - sz)], 1)
            p[fx] = np.stack([np.full(fx.sum(), cx + .5 * sx), cy + (u[fx] - .5) * sy, 0.75 - v[fx] * sz], 1)
            p[fy] = np.stack([cx + (u[fy] - .5) * sx, np.full(fy.sum(), cy - .5 * sy), 0.75 - v[fy] * sz], 1)
        elif kind == 1:    # sphere: upper hemisphere
            r = 0.5 * sx
            th = np.arccos(u)            # polar angle from the up axis, hemisphere
            ph = 2 * np.pi * v
            p = np.stack([cx + r * np.sin(th) * np.cos(ph), cy + r * np.sin(th) * np.sin(ph),
                          0.75 - r - r * np.cos(th)], 1)
        else:              # cylinder: top disc + half the mantle
            r = 0.5 * sx
            top = face == 0
            side = ~top
            rr = r * np.sqrt(u[top])
            p[top] = np.stack([cx + rr * np.cos(2 * np.pi * v[top]), cy + rr * np.sin(2 * np.pi * v[top]),
                               np.full(top.sum(), 0.75 - sz)], 1)
            ang = np.pi * u[side]
            p[side] = np.stack([cx + r * np.cos(ang), cy - r * np.sin(ang), 0.75 - v[side] * sz], 1)
        pts[o:o + m] = p + rng.normal(0, 0.0005, (m, 3))
        col[o:o + m] = base + rng.normal(0, 0.02, (m, 3))
        o += m
    perm = rng.permutation(n)
    pts, col = pts[perm], col[perm]
    nd = int(n * dup_frac)
    if nd > 0:
        src = rng.integers(0, n, nd)
        dst = rng.integers(0, n, nd)
        pts[dst] = pts[src]
        col[dst] = col[src]
    return np.concatenate([pts, np.clip(col, 0, 1)], 1).astype(np.float32)


def uniform_cube(seed, n=4096):
    rng = np.random.default_rng(seed)
    return rng.uniform(0, 1, (n, 6)).astype(np.float32)


def lattice(seed, n=4096, cells=12, pitch=0.02):
    rng = np.random.default_rng(seed)
    ijk = rng.integers(0, cells, (n, 3))
    xyz = ijk.astype(np.float64) * pitch
    col = rng.uniform(0, 1, (n, 3))
    return np.concatenate([xyz, col], 1).astype(np.float32)


def batch(kind, seeds, n):
    gen = {"table": table_scene, "cube": uniform_cube, "lattice": lattice}[kind]
    return np.stack([gen(int(s), n) for s in seeds], 0)


def scores_like_dataset(seed, b, n):
    """pc_score labels ~ tanh(U[0,1]) (scoredataset.py:79-81 applies tanh to the raw antipodal score)."""
    rng = np.random.default_rng(seed)
    return np.tanh(rng.uniform(0, 1, (b, n))).astype(np.float32)


def scene_grasps(seed, points, n_grasps=600, hit_frac=0.6):
    """Synthetic grasp annotations of one scene in the reference's on-disk format (get_regiondataset.py:63-72):
    `frame` (G,4,4) rigid transforms (columns approach / closing axis / minor normal / centre) and `antipodal_score` (G,).
    A fraction of the grasps sits within a few millimetres of cloud points so that centres find a label, the rest
    floats around the scene."""
    rng = np.random.default_rng(seed)
    pts = np.asarray(points)[:, :3]
    q = rng.normal(size=(n_grasps, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    rot = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                    2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                    2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], axis=1).reshape(n_grasps, 3, 3)
    near = rng.random(n_grasps) < hit_frac
    centre = pts[rng.integers(0, len(pts), n_grasps)] + rng.normal(scale=0.002, size=(n_grasps, 3))
    centre[~near] = pts.mean(0) + rng.normal(scale=0.3, size=((~near).sum(), 3))
    frame = np.tile(np.eye(4), (n_grasps, 1, 1))
    frame[:, :3, :3] = rot
    frame[:, :3, 3] = centre
    return {"frame": frame, "antipodal_score": rng.random(n_grasps)}


def write_scene_file(path, seed, points, **kw):
    """Pickle scene_grasps(...) the way the reference's dataset stores a scene (np.load(path, allow_pickle=True) reads it)."""
    import pickle
    with open(path, "wb") as f:
        pickle.dump(scene_grasps(seed, points, **kw), f)
    return path


def training_scene(seed, n_view=30000, n_grasps=3000):
    """One scene in the reference's training-set format (dataset_utils/scoredataset.py:63-66, get_regiondataset.py:63-72,
    eval_utils/torch_scene_point_cloud.py:9-13): the single-view cloud with colours, per-point grasp scores and object
    labels (0 = table), the grasp annotations, and a scene cloud with normals for the evaluation code."""
    rng = np.random.default_rng(seed)
    cloud = table_scene(seed, n_view, dup_frac=0.0)
    xyz, rgb = cloud[:, :3].astype(np.float64), cloud[:, 3:6].astype(np.float64)
    label = (xyz[:, 2] > 0.7515).astype(np.float64) * (1 + (np.abs(np.floor(xyz[:, 0] * 10)) % 5))   # 0 = table plane
    score = np.where(label > 0, rng.random(n_view) * 2.0, rng.random(n_view) * 0.2)                    # tanh'ed by the dataset
    d = scene_grasps(seed + 1, cloud, n_grasps=n_grasps, hit_frac=0.9)
    normals = np.tile(np.array([0.0, 0.0, 1.0]), (n_view, 1))
    d.update(view_cloud=xyz, view_cloud_color=rgb, view_cloud_score=score, view_cloud_label=label, scene_cloud=xyz,
             scene_normal=normals)
    return d


def write_dataset(root, n_scenes=10, seed=0, split="training_data", **kw):
    """<root>/<split>/scene_XXXX.p files that dataset_utils.scoredataset.ScoreDataset(all_points_num, root, tag, ...)
    lists and np.load(..., allow_pickle=True)'s (scoredataset.py:37-62): split "training_data" feeds the train / validate
    tags (80 / 20 %), "training_data_test" the test tag."""
    import os
    import pickle
    out = os.path.join(root, split)
    os.makedirs(out, exist_ok=True)
    paths = []
    for i in range(n_scenes):
        path = os.path.join(out, f"scene_{i:04d}.p")
        with open(path, "wb") as f:
            pickle.dump(training_scene(seed + 10 * i, **kw), f)
        paths.append(path)
    return paths
