"""Stand-in for the `transforms3d` package: the three functions the reference calls (utils.py:436-437,
dataset_utils/eval_score/eval_utils/evaluation_data_generator.py:43).  Quaternions are (w, x, y, z), like transforms3d."""
from . import euler, quaternions  # noqa: F401
