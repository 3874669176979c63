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
from regnet_for_3d_grasping_b200 import sharding, synth, weights  # noqa: E402
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
    torch.manual_seed(0)
    net = ScoreNetwork(training=True).to(dev)
    net.load_state_dict(weights.random_scorenet_state(seed=0))
    net.train()
    model = sharding.wrap_ddp(net, dev) if world > 1 else net
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    pc = torch.from_numpy(synth.batch("table", sharding.shard_seeds(rank, args.batch), args.points)).to(dev)
    tgt = torch.from_numpy(synth.scores_like_dataset(7 + rank, args.batch, args.points)).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        _, _, loss = model(pc, tgt)
        loss = loss.sum()
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = sharding.max_over_ranks([e0.elapsed_time(e1)], dev)[0]
    nparam = sum(p.numel() for p in net.parameters())
    if rank == 0:
        print(json.dumps({"metric": "clouds/s, ScoreNet training step (pretrain_score: fwd + MSE + bwd + all-reduce + Adam)",
                          "value": world * args.batch * args.steps / (ms * 1e-3), "unit": "clouds/s", "n_gpus": world,
                          "ms_per_step": ms / args.steps, "steps": args.steps, "batch_per_gpu": args.batch,
                          "points": args.points, "loss": float(loss), "parameters": nparam,
                          "grad_allreduce_bytes_per_step": 4 * nparam if world > 1 else 0,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "path": "modules.py op-by-op: this repo's point operators, BN+ReLU and max-pool training kernels + torch GEMMs / autograd"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
