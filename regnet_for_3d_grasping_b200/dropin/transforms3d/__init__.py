"""Stand-in for the `transforms3d` package: the three functions the reference calls (utils.py:436-437,
dataset_utils/eval_score/eval_utils/evaluation_data_generator.py:43).  Quaternions are (w, x, y, z), like transforms3d."""
from _bootstrap import installed_elsewhere as _installed_elsewhere   # the drop-in directory is on sys.path (that is how this package was found)

if _installed_elsewhere("transforms3d") is None:      # otherwise sys.modules["transforms3d"] now is the real package
    from . import euler, quaternions  # noqa: F401
