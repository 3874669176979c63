"""Bring-up check of the TMEM-chained SA level-0 kernel (csrc/sa0_chain.cu) against a float64 torch evaluation of the
same chain, with the raw accumulators of layers 0 and 1 dumped by the kernel's debug instance.  One subprocess per
instance (a protocol bug traps the context, which must not take the other run with it).

    python scripts/sa0_chain_check.py            # dump instance (1) and production instance (0), one subprocess each
    python scripts/sa0_chain_check.py 0          # one instance, in-process
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(variant, B=2, N=4096, M=1024):
    import torch
    from regnet_for_3d_grasping_b200 import _lib, pn2_ext, synth
    lib = _lib.load()
    torch.manual_seed(0)
    pc = torch.from_numpy(synth.batch("table", range(B), N)).cuda()
    xyz = pc[:, :, :3].permute(0, 2, 1)
    idx = pn2_ext.farthest_point_sample(xyz, M)
    new_xyz = xyz.gather(2, idx.unsqueeze(1).expand(B, 3, M)).contiguous()
    nbr64, _ = pn2_ext.ball_query(xyz, new_xyz, 0.05, 64)
    nbr = nbr64.to(torch.int32).contiguous()
    g = torch.Generator(device="cpu").manual_seed(1)
    W0 = (torch.randn(128, 6, generator=g) * 0.5).cuda()
    W1 = (torch.randn(128, 128, generator=g) * 0.12).cuda()
    W2 = (torch.randn(256, 128, generator=g) * 0.12).cuda()
    sc = [(torch.rand(n, generator=g) + 0.5) * (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float() for n in (128, 128, 256)]
    sh = [torch.randn(n, generator=g) * 0.3 for n in (128, 128, 256)]
    sc = [t.cuda() for t in sc]
    sh = [t.cuda() for t in sh]
    rows = B * M * 64
    out = torch.full((B * M, 256), -7.0, device="cuda")
    dbg = torch.full((rows, 256), -7.0, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.regnet_sa0_chain(p(pc), p(new_xyz), p(nbr), B, N, M, p(W0), p(sc[0]), p(sh[0]), p(W1), p(sc[1]), p(sh[1]),
                              p(W2), p(sc[2]), p(sh[2]), p(out), p(dbg), variant, None)
    _lib.check(rc)
    torch.cuda.synchronize()
    # float64 reference
    bidx = torch.arange(B, device="cuda").view(B, 1, 1).expand(B, M, 64)
    pts = pc[bidx, nbr.long()]                                   # (B,M,64,6)
    rel = pts[..., :3] - new_xyz.permute(0, 2, 1).unsqueeze(2)   # fp32 subtraction like the kernel
    v = torch.cat([pts[..., 3:6], rel], dim=-1).reshape(rows, 6).double()
    acc0 = v @ W0.double().t()
    y0 = torch.relu(acc0 * sc[0].double() + sh[0].double())
    acc1 = y0 @ W1.double().t()
    y1 = torch.relu(acc1 * sc[1].double() + sh[1].double())
    acc2 = y1 @ W2.double().t()
    y2 = torch.relu(acc2 * sc[2].double() + sh[2].double())
    want = y2.view(B * M, 64, 256).max(dim=1).values

    def err(got, ref):
        return ((got.double() - ref).abs().max() / ref.abs().max()).item()

    e2 = err(out, want)
    if variant == 1:   # the dump instance: raw accumulators of layers 0 and 1
        e0, e1 = err(dbg[:, :128], acc0), err(dbg[:, 128:], acc1)
        print(f"variant {variant}: acc0 rel err {e0:.3e}  acc1 rel err {e1:.3e}  pooled out rel err {e2:.3e}", flush=True)
    else:
        print(f"variant {variant}: pooled out rel err {e2:.3e}", flush=True)
    return e2


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    if len(sys.argv) > 1:
        run(int(sys.argv[1]))
    else:
        for v in (1, 0):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(v)], capture_output=True, text=True, timeout=300)
            tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
            print(f"[variant {v}] rc={r.returncode}: " + " | ".join(tail), flush=True)
