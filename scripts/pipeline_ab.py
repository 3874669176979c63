"""A/B of the stream layouts of the ScoreNet plan at the BASELINE batch: side-stream mode 0/1/2, with and without
cross-step prefetch.  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan  # noqa: E402

B, N, K = 15, 25600, 12
pc = torch.from_numpy(synth.batch("table", range(B), N)).cuda()
pcs = [pc, pc.clone()]
sd = weights.random_scorenet_state(seed=0)
feat = torch.empty(B, N, 256, device="cuda")
score = torch.empty(B, N, device="cuda")
for mode in (0, 1, 2):
    plan = ScoreNetPlan(B, N, "cuda", side_stream=mode)
    plan.bind_state(sd)
    for prefetch in (False, True):
        def run(n):
            if prefetch:
                plan.prefetch(pcs[0])
            for i in range(n):
                if prefetch and i + 1 < n:
                    plan.prefetch(pcs[(i + 1) & 1])
                plan.forward(pcs[i & 1], feat, score)
        run(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(K)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"side_mode={mode} prefetch={prefetch}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} clouds/s", flush=True)
    plan.close()
