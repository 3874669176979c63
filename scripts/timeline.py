"""Per-launch timeline (start offset + duration) of the pipelined ScoreNet loop, streams NOT serialised.  GPU box only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan  # noqa: E402

B, N = 15, 25600
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pc = torch.from_numpy(synth.batch("table", range(B), N)).cuda()
pcs = [pc, pc.clone()]
plan = ScoreNetPlan(B, N, "cuda", side_stream=mode)
plan.bind_state(weights.random_scorenet_state(seed=0))
feat = torch.empty(B, N, 256, device="cuda")
score = torch.empty(B, N, device="cuda")


def run(n):
    plan.prefetch(pcs[0])
    for i in range(n):
        if i + 1 < n:
            plan.prefetch(pcs[(i + 1) & 1])
        plan.forward(pcs[i & 1], feat, score)


run(3)
torch.cuda.synchronize()
recs = plan.timeline(lambda: run(steps))
print(f"# side_mode={mode} FPS_FORCE={os.environ.get('REGNET_FPS_FORCE')}  label start_ms dur_ms end_ms")
for label, ms, t0 in recs:
    print(f"{label:18s} {t0:9.3f} {ms:8.3f} {t0 + ms:9.3f}")
print("# total", max(t0 + ms for _, ms, t0 in recs))
