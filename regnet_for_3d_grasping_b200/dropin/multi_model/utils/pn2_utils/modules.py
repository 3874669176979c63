"""reference: pn2_utils/modules.py (the five classes REGNet instantiates)"""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200 import function as _F  # noqa: F401
from regnet_for_3d_grasping_b200.modules import (FarthestPointSampler, FeatureInterpolator, PointNetSAModule,  # noqa: F401
                                                 PointnetFPModule, QueryGrouper)
