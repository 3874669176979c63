"""reference: multi_model/utils/pointnet2.py -> PointNet2Seg (ScoreNet body), PointNet2TwoStage / PointNet2Refine heads"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.pointnet2 import PointNet2Seg  # noqa: F401
from regnet_for_3d_grasping_b200.region_heads import PointNet2Refine, PointNet2TwoStage  # noqa: F401
