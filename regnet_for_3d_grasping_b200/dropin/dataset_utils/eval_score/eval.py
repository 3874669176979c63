"""reference: dataset_utils/eval_score/eval.py -> eval_test / eval_validate, batched over the grasps (this repository's
grasp_eval.py) instead of the reference's per-grasp Python loops.  The reference's classes stay reachable through the
extended package path (dataset_utils.eval_score.eval_utils.evaluation_data_generator)."""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.grasp_eval import eval_test, eval_validate  # noqa: F401
