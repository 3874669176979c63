"""BASELINE config 5 shape of a step (train.py --mode train, train.py:227-247) on N GPUs of one node: ScoreNetwork in
train mode on 15 synthetic 25 600-point clouds per GPU, get_grasp_allobj (64 centres per cloud, crops of 256 / 1024 points,
grasp labels looked up in per-scene annotation files), GripperRegionNetwork training call (anchor + refine losses),
backward through both networks, gradient all-reduce (DDP over NCCL), two Adam optimisers.  GPU box only.

    python scripts/train_full_step.py [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_full_step.py

Prints one JSON line on rank 0 (whole-job clouds/s, max-over-ranks CUDA-event time)."""
import argparse
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import region, sharding, synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.gripper_region_network import GripperRegionNetwork  # noqa: E402
from regnet_for_3d_grasping_b200.score_network import ScoreNetwork  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=15)
    ap.add_argument("--points", type=int, default=25600)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    width, height, depth = 0.08, 0.010, 0.06                       # train.py:70-75
    params = [64, 0.5, 256, 0.1, 1024, 0.8, width, height, depth]  # train.py:77-90
    gripper_params = [width, height, depth]
    torch.manual_seed(0)
    score_net = ScoreNetwork(training=True).to(dev)
    score_net.load_state_dict(weights.random_scorenet_state(seed=0))
    score_net.train()
    region_net = GripperRegionNetwork(training=True, group_num=256, gripper_num=64, grasp_score_threshold=0.5, radius=depth,
                                      reg_channel=10).to(dev).train()
    score_model = sharding.wrap_ddp(score_net, dev) if world > 1 else score_net
    region_model = sharding.wrap_ddp(region_net, dev, find_unused_parameters=True) if world > 1 else region_net
    opt_score = torch.optim.Adam(score_net.parameters(), lr=1e-3)
    opt_region = torch.optim.Adam(region_net.parameters(), lr=1e-3)
    seeds = list(sharding.shard_seeds(rank, args.batch))
    host = synth.batch("table", seeds, args.points)
    pc = torch.from_numpy(host).to(dev)
    tgt = torch.from_numpy(synth.scores_like_dataset(7 + rank, args.batch, args.points)).to(dev)
    tmp = tempfile.mkdtemp(prefix="regnet_scenes_")
    paths = [synth.write_scene_file(os.path.join(tmp, f"scene{b}.p"), 500 + seeds[b], host[b], n_grasps=3000, hit_frac=0.9)
             for b in range(args.batch)]
    stats = {}

    def step(i):
        opt_score.zero_grad(set_to_none=True)
        opt_region.zero_grad(set_to_none=True)
        all_feature, output_score, loss = score_model(pc, tgt)
        (center_pc, center_idx, gi, gp, gmi, gmp, labels) = region.get_grasp_allobj(pc, output_score.detach(), params, paths,
                                                                                   seed=100 + i)
        out = region_model(gp, gmp, gi, gmi, center_pc, center_idx, pc, all_feature, gripper_params, labels, paths)
        total = loss.sum() + out[3][0].sum()
        if out[13][0] is not None:
            total = total + out[13][0].sum()
        total.backward()
        opt_score.step()
        opt_region.step()
        stats.update(loss=float(total.detach()), labelled=int((labels[:, :, 7] != -1).sum()), refined=0 if out[11] is None else len(out[11]))

    for i in range(args.warmup):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks([e0.elapsed_time(e1)], dev)[0]
    # one extra, synchronised step for a wall-clock breakdown of the stages (diagnostic, not part of the timing)
    import time
    br = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        br[name] = round((time.perf_counter() - t0) * 1e3, 2)
        return time.perf_counter()

    opt_score.zero_grad(set_to_none=True)
    opt_region.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    all_feature, output_score, loss = score_model(pc, tgt)
    t = tick("scorenet_forward_ms", t)
    got = region.get_grasp_allobj(pc, output_score.detach(), params, [], seed=1)
    t = tick("centres_and_crops_ms", t)
    labels = region.get_center_grasp(got[1], got[0], paths, depth)
    t = tick("label_lookup_ms", t)
    out = region_model(got[3], got[5], got[2], got[4], got[0], got[1], pc, all_feature, gripper_params, labels, paths)
    t = tick("region_forward_and_losses_ms", t)
    total = loss.sum() + out[3][0].sum() + (out[13][0].sum() if out[13][0] is not None else 0.0)
    total.backward()
    t = tick("backward_ms", t)
    opt_score.step()
    opt_region.step()
    t = tick("optimisers_ms", t)
    if rank == 0:
        print(json.dumps({"metric": "clouds/s, full REGNet training step (ScoreNet + region crops + labels + GraspRegionNet + "
                                    "RefineNet losses, bwd, all-reduce, Adam)",
                          "value": world * args.batch * args.steps / (ms * 1e-3), "unit": "clouds/s", "n_gpus": world,
                          "ms_per_step": ms / args.steps, "steps": args.steps, "batch_per_gpu": args.batch,
                          "points": args.points, "centres_per_cloud": 64, "labelled_centres_last_step": stats["labelled"],
                          "refined_grasps_last_step": stats["refined"], "loss": stats["loss"],
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30, "breakdown_synchronised": br}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
