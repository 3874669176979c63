"""Fused-operand GEMM (gemm_fused_a.cu) bring-up: values against the materialised path and serial per-kernel times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regnet_for_3d_grasping_b200 import synth, weights
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan

B, N = 15, 25600
sd = weights.random_scorenet_state(seed=8)
pc = torch.from_numpy(synth.batch("table", range(400, 400 + B), N)).cuda()
res = {}
for fused in (0, 1):
    plan = ScoreNetPlan(B, N, "cuda")
    plan.set_option("sa_fused_a", fused)
    plan.bind_state(sd)
    f, s = plan.forward(pc)
    torch.cuda.synchronize()
    res[fused] = (f.clone(), s.clone())
    for _ in range(3):
        prof = plan.profile_forward(pc)
    keep = [(k, round(v, 4)) for k, v in prof if k.startswith(("gemm.sa1", "gemm.sa2", "sa_operand", "sa_fold"))]
    print("fused" if fused else "plain", keep, "sum", round(sum(v for _, v in keep), 4), flush=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for _ in range(5):
        plan.forward(pc)
    ev[0].record()
    for _ in range(20):
        plan.forward(pc)
    ev[1].record()
    torch.cuda.synchronize()
    print("  unpipelined forward ms", ev[0].elapsed_time(ev[1]) / 20, flush=True)
    plan.close()
f0, s0 = res[0]
f1, s1 = res[1]
print("max |df| / max|f| =", float((f1 - f0).abs().max()) / float(f0.abs().max()), " max |ds| =", float((s1 - s0).abs().max()))
