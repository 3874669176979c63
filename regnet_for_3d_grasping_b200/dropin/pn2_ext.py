"""Top-level `pn2_ext` for the reference's `import pn2_ext` (pn2_utils/function.py:2).
Put this directory on PYTHONPATH."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from regnet_for_3d_grasping_b200.pn2_ext import *  # noqa: F401,F403,E402
from regnet_for_3d_grasping_b200.pn2_ext import (ball_query, farthest_point_sample, group_points_backward,  # noqa: F401,E402
                                                 group_points_forward, interpolate_backward, interpolate_forward,
                                                 point_search)
