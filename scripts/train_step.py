"""BASELINE config 4 shape of a step (train.py --mode pretrain_score, train.py:143-149) on N GPUs of one node:
ScoreNetwork in train mode on 15 synthetic 25 600-point clouds per GPU -- forward (this repo's point operators under
torch's convolutions / batch-norm, `modules.py` path), MSE loss, backward (scatter-add kernels A4 / A7), gradient
all-reduce (DDP over NCCL, the only exchange of the path), Adam.  GPU box only.

    python scripts/train_step.py [--steps 5] [--batch 15]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_step.py

Prints one JSON line on rank 0: whole-job clouds/s (max-over-ranks CUDA-event time), peak memory, gradient bytes."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


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
    stepper = ts.ScoreTrainStep(dev, rank, world, args.batch, args.points)
    net = stepper.net
    ms = ts.time_steps(stepper, args.steps, args.warmup, dev)
    loss = stepper.last_loss
    nparam = sum(p.numel() for p in net.parameters())
    ar_ms = ts.time_allreduce(stepper.grad_bytes, dev)      # collective: every rank
    if rank == 0:
        print(json.dumps({"metric": "clouds/s, ScoreNet training step (pretrain_score: fwd + MSE + bwd + all-reduce + Adam)",
                          "value": world * args.batch * args.steps / (ms * 1e-3), "unit": "clouds/s", "n_gpus": world,
                          "ms_per_step": ms / args.steps, "steps": args.steps, "batch_per_gpu": args.batch,
                          "points": args.points, "loss": float(loss), "parameters": nparam,
                          "grad_allreduce_bytes_per_step": 4 * nparam if world > 1 else 0,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "allreduce_ms_isolated": ar_ms, "path": "modules.py: this repo's point operators + chained shared MLPs on the tcgen05 engine (conv_train.py), one flat gradient all-reduce"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
