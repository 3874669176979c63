"""Generate golden vectors from the REAL reference CUDA kernels (oracle/_ref/pn2_ext_ref.so) on a GPU box.

TEST INFRASTRUCTURE ONLY.  Run on the B200 box:  python oracle/gen_golden_gpu.py gpurun_out/golden
then copy gpurun_out/golden/ref_cuda_ops.npz to tests/golden/ and commit it.  This is what pins the CPU oracle
(oracle/pn2_oracle.c): tests/test_oracle_golden.py replays the same seeded inputs through the oracle and
requires bit-equal indices / counts / squared distances.

The inputs are NOT stored: `cases()` below regenerates them from seeds (numpy default_rng + synth), so the
tests and this script share one definition.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from regnet_for_3d_grasping_b200 import synth  # noqa: E402


def cases():
    """name -> dict(kind, pts (B,N,6) fp32, M, radius, K).  Shared by the generator and the tests."""
    c = {}
    c["c1_table_4096"] = dict(pts=synth.batch("table", [0], 4096), M=1024, radius=0.05, K=64)
    c["c1_cube_4096"] = dict(pts=synth.batch("cube", [1], 4096), M=1024, radius=0.1, K=64)
    c["c1_lattice_4096"] = dict(pts=synth.batch("lattice", [2], 4096), M=1024, radius=0.05, K=64)
    c["table_25600_b2"] = dict(pts=synth.batch("table", [3, 4], 25600), M=5120, radius=0.02, K=64)
    c["lattice_5120_b3"] = dict(pts=synth.batch("lattice", [5, 6, 7], 5120), M=1024, radius=0.08, K=64)
    c["cube_1024_b2"] = dict(pts=synth.batch("cube", [8, 9], 1024), M=256, radius=0.32, K=64)
    c["cube_100"] = dict(pts=synth.batch("cube", [10], 100), M=37, radius=0.3, K=8)      # block 128, ragged
    c["cube_20_all"] = dict(pts=synth.batch("cube", [11], 20), M=20, radius=0.5, K=4)    # M == N, block 32
    c["cube_9"] = dict(pts=synth.batch("cube", [12], 9), M=5, radius=0.4, K=3)           # block 16
    c["cube_5"] = dict(pts=synth.batch("cube", [13], 5), M=3, radius=0.01, K=2)          # default case, empty balls
    dup = synth.batch("cube", [14], 600)
    dup[0, 300:] = dup[0, :300]                                                           # every point twice
    c["dup_600"] = dict(pts=dup, M=400, radius=0.2, K=16)                                 # forces zero-distance picks
    same = np.repeat(synth.batch("cube", [15], 1)[:, :1], 64, axis=1)
    c["allsame_64"] = dict(pts=same.copy(), M=8, radius=0.1, K=4)                         # all distances 0
    return c


def run_reference(ext, torch, out_path, dev="cuda"):
    out = {}
    for name, c in cases().items():
        pc = torch.from_numpy(c["pts"]).to(dev)
        xyz = pc[:, :, :3].permute(0, 2, 1)                 # (B,3,N) non-contiguous view, like score_network.py:46
        B, _, N = xyz.shape
        idx = ext.farthest_point_sample(xyz, c["M"])
        new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, c["M"]))
        bq, cnt = ext.ball_query(xyz, new_xyz, c["radius"], c["K"])
        out[name + ".fps"] = idx.cpu().numpy().astype(np.int32)
        out[name + ".bqcnt"] = cnt.cpu().numpy().astype(np.int32)
        bqn = bq.cpu().numpy()
        if bqn.size <= 300000:
            out[name + ".bq"] = bqn.astype(np.int32)
        else:
            out[name + ".bq_head"] = bqn[:, :256].astype(np.int32)
            w = (np.arange(c["K"], dtype=np.int64) * 2654435761 % 1000003 + 1)
            out[name + ".bq_rowhash"] = (bqn * w).sum(-1)
        if c["M"] >= 3:
            nn, nnd = ext.point_search(xyz, new_xyz, 3)
            nnn, nndn = nn.cpu().numpy(), nnd.cpu().numpy()
            if nnn.size <= 300000:
                out[name + ".nn"] = nnn.astype(np.int32)
                out[name + ".nnd"] = nndn
            else:
                out[name + ".nn_head"] = nnn[:, :4096].astype(np.int32)
                out[name + ".nnd_head"] = nndn[:, :4096]
                out[name + ".nn_rowhash"] = (nnn * np.array([1, 1000003, 998244353], dtype=np.int64)).sum(-1)
                out[name + ".nnd_sum"] = nndn.astype(np.float64).sum(-1)
    # float ops: group / interpolate forward + backward on one mid-size case
    g = torch.Generator().manual_seed(123)
    c = cases()["cube_1024_b2"]
    pc = torch.from_numpy(c["pts"]).to(dev)
    xyz = pc[:, :, :3].permute(0, 2, 1)
    feat = torch.randn(2, 19, 1024, generator=g).to(dev)
    idx = ext.farthest_point_sample(xyz, 256)
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(2, 3, 256))
    bq, _ = ext.ball_query(xyz, new_xyz, 0.12, 16)
    grouped = ext.group_points_forward(feat, bq)
    gout = torch.randn(2, 19, 256, 16, generator=g).to(dev)
    ggrad = ext.group_points_backward(gout, bq, 1024)
    nn, nnd = ext.point_search(xyz, new_xyz, 3)
    inv = 1.0 / torch.clamp(nnd, min=1e-10)
    w = inv / inv.sum(2, keepdim=True)
    sfeat = torch.randn(2, 19, 256, generator=g).to(dev)
    interp = ext.interpolate_forward(sfeat, nn, w)
    iout = torch.randn(2, 19, 1024, generator=g).to(dev)
    igrad = ext.interpolate_backward(iout, nn, w, 256)
    out.update({"float.grouped": grouped.cpu().numpy(), "float.group_grad": ggrad.cpu().numpy(),
                "float.weight": w.cpu().numpy(), "float.interp": interp.cpu().numpy(),
                "float.interp_grad": igrad.cpu().numpy()})
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    np.savez_compressed(out_path, **out)
    return out


def compare_oracle(out, torch):
    """Replay through the CPU oracle and report mismatches (also done by tests/test_oracle_golden.py)."""
    from oracle import pn2_oracle as O
    bad = 0
    for name, c in cases().items():
        pc = torch.from_numpy(c["pts"])
        xyz = pc[:, :, :3].permute(0, 2, 1)
        B = xyz.shape[0]
        idx = O.farthest_point_sample(xyz, c["M"])
        ok_fps = np.array_equal(idx.numpy(), out[name + ".fps"])
        new_xyz = xyz.gather(2, torch.from_numpy(out[name + ".fps"]).long().unsqueeze(1).expand(B, 3, c["M"]))
        bq, cnt = O.ball_query(xyz, new_xyz, c["radius"], c["K"])
        ok_cnt = np.array_equal(cnt.numpy(), out[name + ".bqcnt"])
        if name + ".bq" in out:
            ok_bq = np.array_equal(bq.numpy(), out[name + ".bq"])
        else:
            w = (np.arange(c["K"], dtype=np.int64) * 2654435761 % 1000003 + 1)
            ok_bq = np.array_equal((bq.numpy() * w).sum(-1), out[name + ".bq_rowhash"])
        ok_nn = True
        if c["M"] >= 3:
            nn, nnd = O.point_search(xyz, new_xyz, 3)
            if name + ".nn" in out:
                ok_nn = np.array_equal(nn.numpy(), out[name + ".nn"]) and np.array_equal(nnd.numpy(), out[name + ".nnd"])
            else:
                ok_nn = np.array_equal(nn.numpy()[:, :4096], out[name + ".nn_head"]) and \
                    np.array_equal(nnd.numpy()[:, :4096], out[name + ".nnd_head"]) and \
                    np.array_equal((nn.numpy() * np.array([1, 1000003, 998244353], dtype=np.int64)).sum(-1), out[name + ".nn_rowhash"])
        print(f"[oracle-vs-refcuda] {name:18s} fps={ok_fps} bqcnt={ok_cnt} bq={ok_bq} nn={ok_nn}", flush=True)
        bad += (not ok_fps) + (not ok_cnt) + (not ok_bq) + (not ok_nn)
    return bad


if __name__ == "__main__":
    import torch
    from oracle import build_ref
    out_dir = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden"
    ext = build_ref.load()
    out = run_reference(ext, torch, os.path.join(out_dir, "ref_cuda_ops.npz"))
    bad = compare_oracle(out, torch)
    print("oracle mismatches:", bad)
    sys.exit(1 if bad else 0)
