import copy, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regnet_for_3d_grasping_b200 import synth, function as F_
from regnet_for_3d_grasping_b200.modules import PointNetSAModule
torch.manual_seed(1)
pts = torch.from_numpy(synth.batch("table", [5, 6], 2048)).cuda()
xyz = pts[:, :, :3].permute(0, 2, 1)
feat0 = torch.randn(2, 64, 2048, device="cuda")
sa = PointNetSAModule(64, (96, 128), 256, 0.1, 64, use_xyz=True).cuda().train()
with torch.no_grad():
    new_xyz = F_.gather_points(xyz, sa.sampler(xyz))
    index, _ = F_.ball_query(xyz, new_xyz, 0.1, 64)
gen = torch.Generator(device="cuda").manual_seed(5)
dy = torch.randn(2, 128, 256, device="cuda", generator=gen)
res = {}
for name, env, dbl in (("linear", "1", False), ("grouped", "0", False), ("f64", "0", True)):
    os.environ["REGNET_TRAIN_LINEAR_FIRST"] = env
    m = copy.deepcopy(sa)
    f = feat0.clone()
    x = xyz
    if dbl:
        m = m.double(); f = f.double(); x = xyz.double()
    f.requires_grad_(True)
    _, out = m(x, f, geometry=(new_xyz.double() if dbl else new_xyz, index))
    (out * (dy.double() if dbl else dy)).sum().backward()
    res[name] = (out.detach().double(), f.grad.double(), m.mlp[0].conv.weight.grad.double().view(96, 67), m.mlp[1].conv.weight.grad.double().view(128, 96))
ref = res["f64"]
for name in ("linear", "grouped"):
    r = res[name]
    def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
    print(name, "out", rel(r[0], ref[0]), "dfeat", rel(r[1], ref[1]), "dW0 xyz cols", rel(r[2][:, :3], ref[2][:, :3]),
          "dW0 feat cols", rel(r[2][:, 3:], ref[2][:, 3:]), "dW1", rel(r[3], ref[3]),
          "| max dW0 xyz", ref[2][:, :3].abs().max().item(), "max dW0 feat", ref[2][:, 3:].abs().max().item())
