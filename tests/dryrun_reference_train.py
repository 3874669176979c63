"""Build-container check that the reference's OWN train.py runs on top of this repository's drop-in surfaces:

    python tests/dryrun_reference_train.py [pretrain_score]

runs /root/reference/train.py --mode pretrain_score (unmodified file, executed with runpy) for one epoch on a synthetic
data set, with `dropin/` ahead of the reference on sys.path, so that `multi_model.*`, `pn2_ext`,
`dataset_utils.get_regiondataset`, `dataset_utils.eval_score.eval`, `open3d`, `transforms3d` and `tensorboardX` resolve to
this repository.  There is no GPU in the build container, so this is a CPU dry run: the CUDA operators are replaced by
the CPU oracle (test infrastructure, as in tests/test_host_logic.py) and the three places where the reference insists on
a CUDA device are neutralised (torch.cuda.set_device, Tensor/Module.cuda, utils.map_model's .to("cuda:N")).  What it
proves: argument parsing, data set + loader, model construction through utils.construct_net, the train / validate loops,
the logger and torch.save(model) all work against the drop-in modules with the reference's file untouched."""
import os
import runpy
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repository root
REF = "/root/reference"


def main(mode="pretrain_score"):
    if not os.path.isdir(REF):
        raise SystemExit("the reference is not available here (build container only)")
    sys.path.insert(0, ROOT)
    from oracle import pn2_oracle
    from regnet_for_3d_grasping_b200 import function, synth
    pn2_oracle.build()
    function.pn2_ext = pn2_oracle.as_pn2_ext()                 # CPU stand-in for the kernels (dry run only)
    tmp = tempfile.mkdtemp(prefix="regnet_dryrun_")
    synth.write_dataset(tmp, n_scenes=5, seed=0, n_view=26000, n_grasps=200)
    synth.write_dataset(tmp, n_scenes=1, seed=900, split="training_data_test", n_view=26000, n_grasps=200)
    for d in ("models", "log"):
        os.makedirs(os.path.join(tmp, d), exist_ok=True)
    sys.path[:0] = [os.path.join(ROOT, "regnet_for_3d_grasping_b200", "dropin"), REF]
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import utils                                               # the reference's utils.py
    utils.map_model = lambda score_model, region_model, *a, **k: (score_model, region_model)
    real_loader = utils.get_dataloader
    utils.get_dataloader = lambda ds, bs, shuffle=True, **k: real_loader(ds, bs, shuffle=shuffle, num_workers=0, pin_memory=False)
    sys.argv = ["train.py", "--mode", mode, "--epoch", "1", "--batch-size", "2", "--gpu", "0", "--gpu-num", "1",
                "--data-path", tmp, "--model-path", os.path.join(tmp, "models") + "/", "--log-path", os.path.join(tmp, "log") + "/",
                "--tag", "dryrun"]
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    runpy.run_path(os.path.join(REF, "train.py"), run_name="__main__")
    saved = os.path.join(tmp, "models", "dryrun", "score_0.model")
    assert os.path.exists(saved), "train.py did not save the model"
    model = torch.load(saved, weights_only=False)
    print(f"dry run ok: {mode} epoch finished in {time.time() - t0:.0f} s; saved model class "
          f"{type(model).__module__}.{type(model).__name__}; data in {tmp}")


if __name__ == "__main__":
    main(*sys.argv[1:2])
