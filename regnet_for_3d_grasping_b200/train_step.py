"""The two training steps BASELINE.json names, as reusable objects (used by bench.py's `train` record and by
scripts/train_step.py / scripts/train_full_step.py):

  ScoreTrainStep   config 4, `train.py --mode pretrain_score` (train.py:143-149): ScoreNetwork forward in train mode,
                   MSE loss, backward, gradient all-reduce, Adam
  FullTrainStep    config 5, `train.py --mode train` (train.py:227-247): the above + get_grasp_allobj (centres, crops,
                   label lookup) + GripperRegionNetwork training call (anchor + refine losses), two Adam optimisers

Multi-GPU: one process per GPU, every rank its own 15 clouds; the only exchange is ONE all-reduce of a flat gradient
buffer per network (sharding.FlatGrads) -- the gradients of all parameters are views into that buffer, so there is no
bucket copy and no reducer graph walk.  BatchNorm statistics stay per replica, like the reference's nn.DataParallel
(utils.py:129-133)."""
import os
import tempfile

import torch

from . import region, sharding, synth, weights
from .gripper_region_network import GripperRegionNetwork
from .score_network import ScoreNetwork

WIDTH, HEIGHT, DEPTH = 0.08, 0.010, 0.06                        # train.py:70-75
REGION_PARAMS = [64, 0.5, 256, 0.1, 1024, 0.8, WIDTH, HEIGHT, DEPTH]   # train.py:77-90
GRIPPER_PARAMS = [WIDTH, HEIGHT, DEPTH]


class ScoreTrainStep:
    def __init__(self, device, rank=0, world=1, batch=15, points=25600, lr=1e-3):
        self.dev, self.rank, self.world, self.batch, self.points = device, rank, world, batch, points
        torch.manual_seed(0)
        self.net = ScoreNetwork(training=True).to(device)
        self.net.load_state_dict(weights.random_scorenet_state(seed=0))
        self.net.train()
        self.grads = sharding.FlatGrads(self.net.parameters())
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr)
        self.seeds = list(sharding.shard_seeds(rank, batch))
        self.host = synth.batch("table", self.seeds, points)
        self.pc = torch.from_numpy(self.host).to(device)
        self.tgt = torch.from_numpy(synth.scores_like_dataset(7 + rank, batch, points)).to(device)
        self.last_loss = None

    @property
    def grad_bytes(self):
        return self.grads.nbytes

    def next_batch(self):
        """The batch the NEXT step will train on (a data loader's look-ahead; here the synthetic batch itself)."""
        return self.pc

    def step(self, i=0):
        self.grads.zero()
        _, _, loss = self.net(self.pc, self.tgt)
        self.net.prefetch(self.next_batch())      # geometry chain of the next step, overlapped with this step's backward
        loss = loss.sum()
        loss.backward()
        self.grads.all_reduce()
        self.opt.step()
        self.last_loss = loss.detach()
        return self.last_loss


class FullTrainStep(ScoreTrainStep):
    def __init__(self, device, rank=0, world=1, batch=15, points=25600, lr=1e-3, n_grasps=3000):
        super().__init__(device, rank, world, batch, points, lr)
        self.region_net = GripperRegionNetwork(training=True, group_num=256, gripper_num=64, grasp_score_threshold=0.5,
                                               radius=DEPTH, reg_channel=10).to(device).train()
        self.region_grads = sharding.FlatGrads(self.region_net.parameters())
        self.opt_region = torch.optim.Adam(self.region_net.parameters(), lr=lr)
        self.tmp = tempfile.mkdtemp(prefix="regnet_scenes_")
        self.paths = [synth.write_scene_file(os.path.join(self.tmp, f"scene{b}.p"), 500 + self.seeds[b], self.host[b],
                                             n_grasps=n_grasps, hit_frac=0.9) for b in range(batch)]
        self.stats = {}

    @property
    def grad_bytes(self):
        return self.grads.nbytes + self.region_grads.nbytes

    def step(self, i=0):
        self.grads.zero()
        self.region_grads.zero()
        all_feature, output_score, loss = self.net(self.pc, self.tgt)
        self.net.prefetch(self.next_batch())
        (center_pc, center_idx, gi, gp, gmi, gmp, labels) = region.get_grasp_allobj(
            self.pc, output_score.detach(), REGION_PARAMS, self.paths, seed=100 + i)
        out = self.region_net(gp, gmp, gi, gmi, center_pc, center_idx, self.pc, all_feature, GRIPPER_PARAMS, labels,
                              self.paths)
        total = loss.sum() + out[3][0].sum()
        if out[13][0] is not None:
            total = total + out[13][0].sum()
        total.backward()
        self.grads.all_reduce()
        self.region_grads.all_reduce()
        self.opt.step()
        self.opt_region.step()
        self.last_loss = total.detach()
        self.stats = dict(labelled=int((labels[:, :, 7] != -1).sum()), refined=0 if out[11] is None else len(out[11]))
        return self.last_loss


def time_steps(stepper, steps, warmup, device):
    """CUDA-event time of `steps` steps after `warmup`, barrier + synchronize on both sides, max over ranks (ms)."""
    import torch.distributed as dist
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    for i in range(warmup):
        stepper.step(i)
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        stepper.step(warmup + i)
    e1.record()
    if multi:
        dist.barrier()
    torch.cuda.synchronize()
    return sharding.max_over_ranks([e0.elapsed_time(e1)], device)[0]


def time_allreduce(nbytes, device, reps=20):
    """Isolated all-reduce of `nbytes` of fp32 (the gradient exchange alone), ms per call, max over ranks; 0 at one GPU."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0.0
    buf = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
    for _ in range(3):
        dist.all_reduce(buf)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    return sharding.max_over_ranks([e0.elapsed_time(e1) / reps], device)[0]
