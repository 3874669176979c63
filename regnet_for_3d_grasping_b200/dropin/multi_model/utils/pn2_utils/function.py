"""reference: pn2_utils/function.py (operator API)"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.function import (ball_query, farthest_point_sample, feature_interpolate,  # noqa: F401
                                                  gather_points, group_points, search_nn_distance)
