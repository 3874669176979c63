"""Time every launch shape of the cluster FPS kernel on the three ScoreNet levels (B=15).  GPU box only."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import _lib, synth  # noqa: E402

lib = _lib.load()
B = 15
for n, m in ((25600, 5120), (5120, 1024), (1024, 256)):
    pts = torch.from_numpy(synth.batch("table", range(B), n)).cuda()
    xyz = pts[:, :, :3].permute(0, 2, 1)
    idx = torch.empty(B, m, dtype=torch.int32, device="cuda")
    nx = torch.empty(B, 3, m, device="cuda")
    ref = None
    for cs in (1, 2, 4, 8):
        for th in (128, 256, 512, 1024):
            def run():
                return lib.regnet_farthest_point_sample_ex(ctypes.c_void_p(xyz.data_ptr()), xyz.stride(0), xyz.stride(1),
                                                           xyz.stride(2), B, n, m, None, ctypes.c_void_p(idx.data_ptr()),
                                                           ctypes.c_void_p(nx.data_ptr()), cs, th,
                                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            rc = run()
            if rc != 0:
                print(f"N={n} M={m} cs={cs} th={th}: {lib.regnet_last_error().decode()}")
                continue
            torch.cuda.synchronize()
            if ref is None:
                ref = idx.clone()
            same = torch.equal(ref, idx)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"N={n} M={m} cs={cs} th={th}: {ms:.3f} ms  {1e3 * ms / (m - 1):.3f} us/iter  "
                  f"{B * n * (m - 1) / ms / 1e6:.1f} Gpts/s  consistent={same}", flush=True)
