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
from regnet_for_3d_grasping_b200 import region  # noqa: E402


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
    from regnet_for_3d_grasping_b200 import train_step as ts
    stepper = ts.FullTrainStep(dev, rank, world, args.batch, args.points)
    ms = ts.time_steps(stepper, args.steps, args.warmup, dev)
    ar_ms = ts.time_allreduce(stepper.grad_bytes, dev)
    stats = dict(stepper.stats, loss=float(stepper.last_loss))
    score_model, region_model, pc, tgt, paths = stepper.net, stepper.region_net, stepper.pc, stepper.tgt, stepper.paths
    opt_score, opt_region = stepper.opt, stepper.opt_region
    params, gripper_params, depth = ts.REGION_PARAMS, ts.GRIPPER_PARAMS, ts.DEPTH
    # one extra, synchronised step for a wall-clock breakdown of the stages (diagnostic, not part of the timing)
    import time
    br = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        br[name] = round((time.perf_counter() - t0) * 1e3, 2)
        return time.perf_counter()

    stepper.grads.zero()
    stepper.region_grads.zero()
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
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "grad_allreduce_bytes_per_step": stepper.grad_bytes if world > 1 else 0,
                          "allreduce_ms_isolated": ar_ms, "breakdown_synchronised": br}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
