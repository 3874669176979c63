"""gemm_fused_a.cu bring-up: serial time of the fused layers under the experiment variants."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regnet_for_3d_grasping_b200 import synth, weights
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan
B, N = 15, 25600
sd = weights.random_scorenet_state(seed=8)
pc = torch.from_numpy(synth.batch("table", range(400, 400 + B), N)).cuda()
plan = ScoreNetPlan(B, N, "cuda")
plan.bind_state(sd)
for v in (0, 1, 2, 3, 4, 7):
    os.environ["REGNET_FUSED_A_VARIANT"] = str(v)
    for _ in range(3):
        prof = plan.profile_forward(pc)
    d = dict(prof)
    print("variant", v, "sa1.l1", round(d["gemm.sa1.l1"], 4), "sa2.l1", round(d["gemm.sa2.l1"], 4), flush=True)
