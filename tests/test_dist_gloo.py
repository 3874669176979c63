"""CPU suite, part 4: the N>1 host logic with gloo, world size 2 (no GPU): shard maps, max-over-ranks timing, and a
DDP training step of the drop-in ScoreNetwork (op-by-op path, oracle ops patched in under the operators) whose
gradients must equal the average of the two ranks' local gradients."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pn2_oracle
        from regnet_for_3d_grasping_b200 import function, sharding, synth, weights
        from regnet_for_3d_grasping_b200.score_network import ScoreNetwork
        function.pn2_ext = pn2_oracle.as_pn2_ext()      # checker's stand-in for the CUDA operators
        res = {}
        res["shards"] = [sharding.shard_range(15, r, world) for r in range(world)]
        res["seeds"] = list(sharding.shard_seeds(rank, 3))
        res["tmax"] = sharding.max_over_ranks([1.0 + rank, 5.0 - rank], "cpu")
        torch.manual_seed(0)
        net = ScoreNetwork(training=True)
        net.load_state_dict(weights.random_scorenet_state(seed=2))
        net.train()
        pc = torch.from_numpy(synth.batch("table", sharding.shard_seeds(rank, 1), 5120))
        tgt = torch.from_numpy(synth.scores_like_dataset(10 + rank, 1, 5120))
        # local gradient first (no DDP), then the DDP step
        torch.manual_seed(123)                        # same dropout mask for both passes
        _, _, loss = net(pc, tgt)
        loss.backward()
        local = net.extrat_featurePN2.conv_score.weight.grad.clone()
        net.zero_grad()
        ddp = sharding.wrap_ddp(net)
        torch.manual_seed(123)
        _, _, loss2 = ddp(pc, tgt)
        loss2.backward()
        res["local"] = local.flatten().tolist()          # plain lists: tensors would travel as shm handles
        res["ddp"] = net.extrat_featurePN2.conv_score.weight.grad.flatten().tolist()
        res["sa_grad_norm"] = net.extrat_featurePN2.sa_modules[0].mlp[0].conv.weight.grad.norm().item()
        # the product's own exchange: gradients as views into one flat buffer, one all-reduce (sharding.FlatGrads)
        net2 = ScoreNetwork(training=True)
        net2.load_state_dict(weights.random_scorenet_state(seed=2))
        net2.train()
        fg = sharding.FlatGrads(net2.parameters())
        fg.zero()
        torch.manual_seed(123)
        _, _, loss3 = net2(pc, tgt)
        loss3.backward()
        g = net2.extrat_featurePN2.conv_score.weight.grad
        res["flat_is_view"] = fg.flat.data_ptr() <= g.data_ptr() < fg.flat.data_ptr() + fg.nbytes
        fg.all_reduce()
        res["flat"] = g.flatten().tolist()
        res["flat_bytes"] = fg.nbytes
        net2.zero_grad(set_to_none=True)            # a caller that drops the views: zero() re-attaches them
        fg.zero()
        res["flat_reattached"] = all(p.grad is not None and float(p.grad.abs().sum()) == 0.0 for p in net2.parameters())
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_sharding_and_ddp(oracle):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=550)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0, r1 = res
    assert r0["shards"] == [(0, 8), (8, 15)]                       # balanced, contiguous, covers the batch
    assert r0["seeds"] == [0, 1, 2] and r1["seeds"] == [1000, 1001, 1002]
    assert r0["tmax"] == [2.0, 5.0] and r1["tmax"] == [2.0, 5.0]  # max over ranks, identical everywhere
    avg = (torch.tensor(r0["local"]) + torch.tensor(r1["local"])) / 2
    assert torch.allclose(torch.tensor(r0["ddp"]), avg, rtol=1e-4, atol=1e-7)   # DDP = average of per-rank gradients
    assert r0["ddp"] == r1["ddp"]                                  # replicas stay in sync
    assert r0["flat_is_view"] and r0["flat_reattached"] and r0["flat_bytes"] == 4 * 5542531
    assert torch.allclose(torch.tensor(r0["flat"]), avg, rtol=1e-4, atol=1e-7) and r0["flat"] == r1["flat"]
    assert r0["sa_grad_norm"] > 0 and abs(r0["sa_grad_norm"] - r1["sa_grad_norm"]) < 1e-6 * max(1.0, r0["sa_grad_norm"])
