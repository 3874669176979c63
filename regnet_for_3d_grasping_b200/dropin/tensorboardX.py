"""Stand-in for the `tensorboardX` package the reference imports (train.py:12, test.py:12, utils.py:8,28): the one class
it uses, SummaryWriter.  Resolves to torch.utils.tensorboard's writer when the `tensorboard` package is importable,
otherwise to a writer that keeps the scalars in memory (add_scalar / add_scalars / add_histogram / flush / close)."""
try:
    from torch.utils.tensorboard import SummaryWriter  # noqa: F401
except Exception:  # tensorboard is not installed: keep the values, write nothing

    class SummaryWriter:
        def __init__(self, logdir=None, *args, **kwargs):
            self.logdir = logdir
            self.scalars = {}

        def add_scalar(self, tag, scalar_value, global_step=None, *args, **kwargs):
            self.scalars.setdefault(tag, []).append((global_step, float(scalar_value)))

        def add_scalars(self, main_tag, tag_scalar_dict, global_step=None, *args, **kwargs):
            for k, v in tag_scalar_dict.items():
                self.add_scalar(f"{main_tag}/{k}", v, global_step)

        def add_histogram(self, *args, **kwargs):
            pass

        def add_text(self, *args, **kwargs):
            pass

        def flush(self):
            pass

        def close(self):
            pass
