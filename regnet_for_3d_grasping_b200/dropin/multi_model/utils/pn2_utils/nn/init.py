"""reference: pn2_utils/nn/init.py"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.nn_layers import init_bn  # noqa: F401
