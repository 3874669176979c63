"""Kernel-level breakdown of one ScoreNet training step (scripts/train_step.py shape) with torch.profiler.  GPU box only."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.score_network import ScoreNetwork  # noqa: E402

B, N = 15, 25600
dev = "cuda"
torch.manual_seed(0)
net = ScoreNetwork(training=True).to(dev)
net.load_state_dict(weights.random_scorenet_state(seed=0))
net.train()
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
pc = torch.from_numpy(synth.batch("table", range(B), N)).to(dev)
tgt = torch.from_numpy(synth.scores_like_dataset(7, B, N)).to(dev)


def step():
    opt.zero_grad(set_to_none=True)
    _, _, loss = net(pc, tgt)
    loss.sum().backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
