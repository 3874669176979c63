"""reference: dataset_utils/get_regiondataset.py -> get_grasp_allobj (region cropping on device)"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.region import get_grasp_allobj, get_group_pc as _get_group_pc  # noqa: F401
from regnet_for_3d_grasping_b200.region import select_score_center as _select_score_center  # noqa: F401
