"""reference: pn2_utils/nn/modules/conv.py"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.nn_layers import Conv1d, Conv2d  # noqa: F401
