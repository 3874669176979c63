"""Where does sa0_chain_kernel spend its time?  Runs the TIMING instance (clock64 around every mbarrier wait) on the
BASELINE-sized level (B=15, N=25600, M=5120) and prints, per warp role, the share of the kernel spent in each wait."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import _lib, pn2_ext, synth  # noqa: E402

B, N, M = 15, 25600, 5120
lib = _lib.load()
pc = torch.from_numpy(synth.batch("table", range(B), N)).cuda()
xyz = pc[:, :, :3].permute(0, 2, 1)
idx = pn2_ext.farthest_point_sample(xyz, M)
new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, M)).contiguous()
nbr = pn2_ext.ball_query(xyz, new_xyz, 0.02, 64)[0].to(torch.int32).contiguous()
g = torch.Generator().manual_seed(1)
W0, W1, W2 = [(torch.randn(o, i, generator=g) * s).cuda() for o, i, s in ((128, 6, .5), (128, 128, .12), (256, 128, .12))]
sc = [(torch.rand(n, generator=g) + 0.5).cuda() for n in (128, 128, 256)]
sh = [(torch.randn(n, generator=g) * 0.3).cuda() for n in (128, 128, 256)]
out = torch.empty(B * M, 256, device="cuda")
timing = torch.zeros(148, 5, 8, dtype=torch.int64, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for variant, buf in ((0, None), (2, timing), (0, None)):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.regnet_sa0_chain(p(pc), p(new_xyz), p(nbr), B, N, M, p(W0), p(sc[0]), p(sh[0]), p(W1), p(sc[1]), p(sh[1]),
                                    p(W2), p(sc[2]), p(sh[2]), p(out), p(buf) if buf is not None else None, variant, None))
    e1.record()
    torch.cuda.synchronize()
    print(f"variant {variant}: call {e0.elapsed_time(e1):.3f} ms (includes weight split + sync)")
t = timing.double().cpu()
tiles = B * M * 64 / 128 / 148
names = {0: ["-", "sempty", "a0empty", "-", "-", "-", "-", "total"],
         1: ["sfull", "a0full", "w", "a1.c0", "a1.c1-3", "a2", "z_empty", "total"],
         2: ["sfull", "acc0/1", "convert", "pool.chunks", "pool.bar", "pool.combine", "acc2", "total"],
         3: ["sfull", "acc0/1", "convert", "pool.chunks", "pool.bar", "pool.combine", "acc2", "total"],
         4: ["sfull", "acc0/1", "convert", "pool.chunks", "pool.bar", "pool.combine", "acc2", "total"]}
roles = ["producer", "mma", "worker0", "worker1", "worker2"]
print(f"tiles per CTA ~{tiles:.0f}; mean over CTAs, cycles per tile:")
for r in range(5):
    m = t[:, r, :].mean(0) / tiles
    print(f"  {roles[r]:9s} " + "  ".join(f"{n}={v:.0f}" for n, v in zip(names[r], m.tolist()) if n != "-"))
