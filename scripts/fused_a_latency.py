"""Single-forward latency (synchronised, median of 7) of the plan under the sa_fused_a modes, and the pipelined step."""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regnet_for_3d_grasping_b200 import synth, weights
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan

B, N = int(os.environ.get("B", 15)), 25600
sd = weights.random_scorenet_state(seed=8)
pcs = [torch.from_numpy(synth.batch("table", range(400 + 20 * k, 400 + 20 * k + B), N)).cuda() for k in range(2)]
for mode in (0, 1, 2, 3):
    plan = ScoreNetPlan(B, N, "cuda")
    plan.set_option("sa_fused_a", mode)
    plan.bind_state(sd)
    names = [label for label, _ in plan.profile_forward(pcs[0])]
    for _ in range(3):
        plan.forward(pcs[0])
    torch.cuda.synchronize()
    lat = []
    for i in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.forward(pcs[i & 1])
        e1.record()
        torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1))
    # pipelined
    plan.prefetch(pcs[0])
    for i in range(10):
        plan.prefetch(pcs[(i + 1) & 1])
        plan.forward(pcs[i & 1])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        plan.prefetch(pcs[(i + 1) & 1])
        plan.forward(pcs[i & 1])
    e1.record()
    torch.cuda.synchronize()
    plan.forward(pcs[0])
    torch.cuda.synchronize()
    print(f"mode {mode}: lone forward {statistics.median(lat):.3f} ms (min {min(lat):.3f}), pipelined {e0.elapsed_time(e1) / 40:.3f} ms/step,"
          f" profile has sa_operand.1: {'sa_operand.1' in names}", flush=True)
    plan.close()
