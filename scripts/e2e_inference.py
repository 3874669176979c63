"""BASELINE.json configs[2]: ScoreNet + GraspRegionNet + RefineNet end-to-end inference on one 25 600-point cloud with the
parameters of the reference's test.py:61-81 (center_num 4000, group_num 256, group_num_more 2048, gripper_num 64) -- the
calls of test.py:134-140 through this repository's drop-in modules, seeded random weights, synthetic table-top cloud.
Prints one JSON line with the per-stage CUDA-event times (median of the timed repeats).  GPU box only."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import region, synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.gripper_region_network import GripperRegionNetwork  # noqa: E402
from regnet_for_3d_grasping_b200.score_network import ScoreNetwork  # noqa: E402

width, height, depth = 0.08, 0.010, 0.06
params = [4000, 0.5, 256, 0.1, 2048, 0.8, width, height, depth]
gripper_params = [width, height, depth]
dev = "cuda"
torch.manual_seed(0)
score_model = ScoreNetwork(training=False).to(dev).eval()
score_model.load_state_dict(weights.random_scorenet_state(seed=0))
region_model = GripperRegionNetwork(training=True, group_num=256, gripper_num=64, grasp_score_threshold=0.5, radius=depth,
                                    reg_channel=10)
region_model.load_state_dict(weights.seeded_state_like(region_model.state_dict(), seed=46))
region_model = region_model.to(dev).eval()
pc = torch.from_numpy(synth.batch("table", [7], 25600)).to(dev)


def once():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with torch.no_grad():
        ev[0].record()
        all_feature, output_score, _ = score_model(pc)
        ev[1].record()
        (center_pc, center_pc_index, pc_group_index, pc_group, pc_group_more_index, pc_group_more, _) = \
            region.get_grasp_allobj(pc, output_score, params, [], True)
        ev[2].record()
        out = region_model(pc_group, pc_group_more, pc_group_index, pc_group_more_index, center_pc, center_pc_index, pc,
                           all_feature, gripper_params, None, [])
        ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)], out


for _ in range(3):
    once()
runs = [once() for _ in range(7)]
t = [statistics.median(r[0][i] for r in runs) for i in range(3)]
out = runs[-1][1]
line = {"workload": "ScoreNet + GraspRegionNet + RefineNet inference, 1 x 25600-pt synthetic cloud, test.py parameters "
                    "(4000 centres, groups of 256 / 2048 points, 64-point closing-box crops)",
        "scorenet_ms": round(t[0], 3), "centres_and_crops_ms": round(t[1], 3), "region_and_refine_ms": round(t[2], 3),
        "total_ms": round(sum(t), 3), "grasps_stage1": int(out[0].shape[0]),
        "accepted_closing_boxes": None if out[11] is None else int(sum(int(k) for k in out[9])),
        "weights": "seeded random", "data": "synthetic"}
print(json.dumps(line))
