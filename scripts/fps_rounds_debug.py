import sys, os, ctypes
sys.path.insert(0, '/root/repo')
import torch
from regnet_for_3d_grasping_b200 import synth, pn2_ext, _lib
lib = _lib.load()
fn = lib.regnet_debug_fps_counters
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_uint * 4)()
for name, B, N, M in (("l0", 15, 25600, 5120), ("b1", 1, 25600, 5120), ("l1", 15, 5120, 1024)):
    pts = torch.from_numpy(synth.batch("table", range(B), N)).cuda()[:, :, :3].permute(0, 2, 1).contiguous()
    fn(out, 1)
    idx = pn2_ext.farthest_point_sample(pts, M)
    torch.cuda.synchronize()
    fn(out, 1)
    print(name, "rounds of cloud 0:", out[0], "picks per round:", (M - 1) / max(out[0], 1))
