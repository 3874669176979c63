"""Second baseline: the reference's OWN CUDA kernels on the same B200, next to this repo's operators.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing in the product path imports this.  Run on the GPU box:

    python oracle/bench_ref_cuda.py > gpurun_out/ref_cuda_baseline.json

What is timed (CUDA events, 2 warm-ups, median of 5):
  * per operator, on the level shapes of BASELINE configs[1] (B = 15) and on configs[0] (1 x 4096):
      farthest_point_sample / ball_query / point_search of oracle/_ref/pn2_ext_ref.so (the reference's .cu files,
      bodies untouched, built by oracle/build_ref.py)  vs  regnet_for_3d_grasping_b200.pn2_ext (same signatures);
      the index outputs are compared bit for bit while we are at it;
  * the whole ScoreNet forward the way the reference computes it on a GPU: oracle/ref_modules.py (the functional
    restatement of multi_model/utils/pointnet2.py, pinned against the real modules by oracle/gen_golden_cpu.py) driving
    the reference CUDA extension and torch's cuDNN convolutions -- once with TF32 convolutions allowed (torch's
    default, i.e. what the reference runs as shipped) and once in strict fp32 -- vs ScoreNetPlan.forward (un-pipelined).
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import build_ref, ref_modules  # noqa: E402
from regnet_for_3d_grasping_b200 import pn2_ext, synth, weights  # noqa: E402
from regnet_for_3d_grasping_b200.scorenet import ScoreNetPlan  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out)


def main():
    ref = build_ref.load()
    dev = "cuda"
    res = {"gpu": torch.cuda.get_device_name(0), "ops": [], "forward": {}}
    shapes = [("C1 1x4096", 1, 4096, 1024, 0.05), ("C2 level 0", 15, 25600, 5120, 0.02),
              ("C2 level 1", 15, 5120, 1024, 0.08), ("C2 level 2", 15, 1024, 256, 0.32)]
    for name, B, N, M, r in shapes:
        pc = torch.from_numpy(synth.batch("table", range(B), N)).to(dev)
        xyz = pc[:, :, :3].permute(0, 2, 1)
        i_ref = ref.farthest_point_sample(xyz, M)
        i_our = pn2_ext.farthest_point_sample(xyz, M)
        new_xyz = xyz.gather(2, i_ref.unsqueeze(1).expand(B, 3, M)).contiguous()
        b_ref, c_ref = ref.ball_query(xyz, new_xyz, r, 64)
        b_our, c_our = pn2_ext.ball_query(xyz, new_xyz, r, 64)
        n_ref, d_ref = ref.point_search(xyz, new_xyz, 3)
        n_our, d_our = pn2_ext.point_search(xyz, new_xyz, 3)
        row = {"shape": name, "B": B, "N": N, "M": M, "radius": r,
               "bit_exact": {"fps": bool(torch.equal(i_ref, i_our)),
                             "ball_query": bool(torch.equal(b_ref, b_our) and torch.equal(c_ref, c_our)),
                             "three_nn": bool(torch.equal(n_ref, n_our) and torch.equal(d_ref, d_our))}}
        for op, f_ref, f_our, work in [
                ("fps", lambda: ref.farthest_point_sample(xyz, M), lambda: pn2_ext.farthest_point_sample(xyz, M),
                 B * N * (M - 1)),
                ("ball_query", lambda: ref.ball_query(xyz, new_xyz, r, 64), lambda: pn2_ext.ball_query(xyz, new_xyz, r, 64),
                 B * M * N),
                ("three_nn", lambda: ref.point_search(xyz, new_xyz, 3), lambda: pn2_ext.point_search(xyz, new_xyz, 3),
                 B * N * M)]:
            t_ref, t_our = timed(f_ref), timed(f_our)
            row[op] = {"reference_cuda_ms": t_ref, "this_repo_ms": t_our, "speedup": t_ref / t_our,
                       "reference_gpts_per_s": work / t_ref / 1e6, "this_repo_gpts_per_s": work / t_our / 1e6}
        res["ops"].append(row)

    # whole forward, BASELINE configs[1]
    B, N = 15, 25600
    pc = torch.from_numpy(synth.batch("table", range(B), N)).to(dev)
    sd = {k: v.to(dev) for k, v in weights.random_scorenet_state(seed=0).items()}
    plan = ScoreNetPlan(B, N, dev)
    plan.bind_state(sd)
    feat, score = plan.forward(pc)
    torch.cuda.synchronize()
    for label, tf32 in (("reference_tf32_default", True), ("reference_fp32_strict", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            rf, rs, _ = ref_modules.scorenet_forward(sd, pc, ref)
            t = timed(lambda: ref_modules.scorenet_forward(sd, pc, ref), warm=1, reps=3)
        err = ((feat - rf).abs().max() / rf.abs().max()).item()
        res["forward"][label] = {"ms_per_batch": t, "clouds_per_s": B / t * 1e3, "feature_rel_err_vs_this_repo": err,
                                 "score_abs_err_vs_this_repo": (score - rs).abs().max().item()}
        del rf, rs
    torch.backends.cudnn.allow_tf32 = True
    t = timed(lambda: plan.forward(pc, feat, score))
    res["forward"]["this_repo_unpipelined"] = {"ms_per_batch": t, "clouds_per_s": B / t * 1e3}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    with torch.no_grad():
        main()
