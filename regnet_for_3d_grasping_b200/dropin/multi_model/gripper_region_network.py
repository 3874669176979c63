"""reference: multi_model/gripper_region_network.py -> GripperRegionNetwork (region + refine stage on device)"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.gripper_region_network import (GripperRegionNetwork, _enumerate_templates,  # noqa: F401
                                                                compute_cos_sim, get_gripper_region_transform)
