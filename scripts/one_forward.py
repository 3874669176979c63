"""Two ScoreNet forwards of the BASELINE batch (B=15 x 25600) through the native plan -- the short command that
ncu wraps (launch list / --set full capture).  argv[1]: engine tc|simt, argv[2]: serial|fork."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import _lib, synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan  # noqa: E402

engine = _lib.ENGINE_SIMT if len(sys.argv) > 1 and sys.argv[1] == "simt" else _lib.ENGINE_TC
fork = not (len(sys.argv) > 2 and sys.argv[2] == "serial")
B, N = 15, 25600
pc = torch.from_numpy(synth.batch("table", range(B), N)).cuda()
plan = ScoreNetPlan(B, N, "cuda", engine=engine, side_stream=fork)
plan.bind_state(weights.random_scorenet_state(seed=0))
for _ in range(2):
    feat, score = plan.forward(pc)
torch.cuda.synchronize()
print("launches per forward:", plan.launch_count, "score mean", score.mean().item())
