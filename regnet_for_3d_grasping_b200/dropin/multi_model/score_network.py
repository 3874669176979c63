"""reference: multi_model/score_network.py -> ScoreNetwork"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.score_network import ScoreNetwork  # noqa: F401
