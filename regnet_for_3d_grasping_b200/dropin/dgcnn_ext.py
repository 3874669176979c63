"""Top-level `dgcnn_ext` for the reference's `import dgcnn_ext` (pn2_utils/functions/gather_knn.py:3)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from regnet_for_3d_grasping_b200.dgcnn_ext import gather_knn_backward, gather_knn_forward  # noqa: F401,E402
