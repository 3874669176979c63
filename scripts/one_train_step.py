"""Two ScoreNet training steps of the BASELINE batch (15 x 25600) -- the short command ncu wraps for the training path
(launch list / --set full capture of conv1x1_tc_kernel and wgrad_tc_kernel)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import train_step as ts  # noqa: E402

torch.cuda.set_device(0)
stepper = ts.ScoreTrainStep(torch.device("cuda", 0))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss = stepper.step(i)
torch.cuda.synchronize()
print("loss", float(loss))
