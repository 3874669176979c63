from .conv import Conv1d, Conv2d  # noqa: F401
from .mlp import SharedMLP  # noqa: F401
