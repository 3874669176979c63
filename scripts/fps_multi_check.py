"""Multi-pick FPS against the one-pick-per-exchange kernel: identical indices, time per call (GPU box)."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    from regnet_for_3d_grasping_b200 import pn2_ext, synth
    out = {}
    for name, B, N, M in (("l0", 15, 25600, 5120), ("l1", 15, 5120, 1024), ("l2", 15, 1024, 256), ("b1", 1, 25600, 5120),
                          ("lattice", 3, 5120, 1024)):
        pts = synth.batch("lattice" if name == "lattice" else "table", range(B), N)
        xyz = torch.from_numpy(pts).cuda()[:, :, :3].permute(0, 2, 1)
        idx = pn2_ext.farthest_point_sample(xyz, M)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            pn2_ext.farthest_point_sample(xyz, M)
        e1.record()
        torch.cuda.synchronize()
        out[name] = (idx.cpu(), e0.elapsed_time(e1) / 5)
    torch.save(out, sys.argv[1])
else:
    res = {}
    for single in ("1", "0"):
        path = f"/tmp/fps_{single}.pt"
        subprocess.check_call([sys.executable, __file__, path], env=dict(os.environ, REGNET_FPS_SINGLE=single))
        res[single] = torch.load(path)
    for name in res["1"]:
        same = torch.equal(res["1"][name][0], res["0"][name][0])
        print(f"{name}: one pick per exchange {res['1'][name][1]:.3f} ms, multi-pick {res['0'][name][1]:.3f} ms, identical indices: {same}")
