"""Acceptance run of SURVEY.md 8(f) row 1 on a GPU box: the reference's OWN train.py and test.py, files untouched,
executed on top of this repository's drop-in surfaces (multi_model.*, pn2_ext, dataset_utils.get_regiondataset,
dataset_utils.eval_score.eval, open3d / transforms3d / tensorboardX stand-ins), real CUDA kernels underneath:

    train.py --mode pretrain_score     one epoch on a synthetic data set in the reference's on-disk format
    train.py --mode train              one epoch of the full REGNet training loop (ScoreNet + region + refine)
    test.py  --folder-name .../test_file/virtual_data    the inference script on the reference's own scene file,
                                       with the two models train.py just saved

The reference tree is NOT part of this repository: copy it to the git-ignored baseline/_ref/REGNet (it travels to the GPU
box with the snapshot) and run   python scripts/run_reference_scripts.py [--ref baseline/_ref/REGNet].
Each script runs in its own process; stdout goes to gpurun_out/ref_<name>.log and one JSON summary line is printed."""
import argparse
import glob
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = r"""
import os, runpy, sys
root, ref, script = sys.argv[1:4]
sys.argv = [script] + sys.argv[4:]
sys.path[:0] = [os.path.join(root, "regnet_for_3d_grasping_b200", "dropin"), ref, root]
runpy.run_path(os.path.join(ref, script), run_name="__main__")
"""


def run(ref, script, argv, log):
    t0 = time.time()
    # torch >= 2.6 unpickles with weights_only=True by default; the reference saves and loads WHOLE modules (torch.save(model),
    # utils.py:63,84), which needs the documented opt-out -- an environment variable, the reference's files stay untouched
    env = dict(os.environ, TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD="1")
    with open(log, "w") as f:
        r = subprocess.run([sys.executable, "-W", "ignore", "-c", RUNNER, ROOT, ref, script] + argv, stdout=f,
                           stderr=subprocess.STDOUT, cwd=ref, timeout=1500, env=env)
    return r.returncode, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.path.join(ROOT, "baseline", "_ref", "REGNet"))
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--batch-size", type=int, default=4)
    args = ap.parse_args()
    ref = os.path.abspath(args.ref)
    if not os.path.exists(os.path.join(ref, "train.py")):
        raise SystemExit(f"no reference checkout at {ref}")
    sys.path.insert(0, ROOT)
    from regnet_for_3d_grasping_b200 import synth
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="regnet_refrun_")
    synth.write_dataset(tmp, n_scenes=args.scenes, seed=0, n_view=26000, n_grasps=400)
    synth.write_dataset(tmp, n_scenes=2, seed=900, split="training_data_test", n_view=26000, n_grasps=400)
    for d in ("models", "log"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    common = ["--gpu", "0", "--gpu-num", "1", "--data-path", tmp, "--model-path", os.path.join(tmp, "models") + "/",
              "--log-path", os.path.join(tmp, "log") + "/"]
    res = {}
    rc, dt = run(ref, "train.py", ["--mode", "pretrain_score", "--epoch", "1", "--batch-size", str(args.batch_size),
                                   "--tag", "accept_score"] + common, os.path.join(out_dir, "ref_train_pretrain_score.log"))
    res["train.py --mode pretrain_score"] = {"rc": rc, "seconds": round(dt, 1),
                                             "saved": sorted(os.path.basename(p) for p in glob.glob(os.path.join(tmp, "models", "accept_score", "*")))}
    rc, dt = run(ref, "train.py", ["--mode", "train", "--epoch", "1", "--batch-size", str(args.batch_size), "--tag", "accept_full"]
                 + common, os.path.join(out_dir, "ref_train_full.log"))
    saved = sorted(os.path.basename(p) for p in glob.glob(os.path.join(tmp, "models", "accept_full", "*")))
    res["train.py --mode train"] = {"rc": rc, "seconds": round(dt, 1), "saved": saved}
    text = open(os.path.join(out_dir, "ref_train_full.log")).read()
    if rc != 0 and "add_eval_log_epoch" in text and "ZeroDivisionError" in text:
        res["train.py --mode train"]["note"] = (
            "the training epoch and every validation batch completed and both models were saved; the reference's epoch "
            "summary (utils.py:376-381) then divides by the number of collision-free predicted grasps, which is 0 for a "
            "network trained for one epoch on 8 synthetic scenes -- a ZeroDivisionError inside the reference's own logger")
    score = os.path.join(tmp, "models", "accept_full", "score_0.model")
    region = os.path.join(tmp, "models", "accept_full", "region_0.model")
    if os.path.exists(score) and os.path.exists(region):
        folder = os.path.join(ref, "test_file", "virtual_data")
        for old in glob.glob(os.path.join(ref, "test_file", "virtual_data_predict", "*")):
            os.remove(old)                       # the checkout ships predictions: only files written by THIS run count
        rc, dt = run(ref, "test.py", ["--gpu", "0", "--gpu-num", "1", "--load-score-path", score, "--load-region-path", region,
                                      "--folder-name", folder, "--model-path", os.path.join(tmp, "models") + "/",
                                      "--log-path", os.path.join(tmp, "log") + "/"], os.path.join(out_dir, "ref_test.log"))
        pred = glob.glob(os.path.join(ref, "test_file", "virtual_data_predict", "*"))
        res["test.py virtual_data"] = {"rc": rc, "seconds": round(dt, 1), "scenes": len(glob.glob(folder + "/*.p")),
                                       "prediction_files": sorted(os.path.basename(p) for p in pred)}
    import torch
    for k, v in res.items():
        log = {"train.py --mode pretrain_score": "ref_train_pretrain_score.log", "train.py --mode train": "ref_train_full.log",
               "test.py virtual_data": "ref_test.log"}[k]
        text = open(os.path.join(out_dir, log)).read()
        v["log_tail"] = text.strip().splitlines()[-3:]
        v["loaded_native_library"] = "libregnet_b200" in text or None
    print(json.dumps({"gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None, "reference": ref,
                      "results": res}, indent=1))


if __name__ == "__main__":
    main()
