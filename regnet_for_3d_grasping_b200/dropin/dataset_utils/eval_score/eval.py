"""reference: dataset_utils/eval_score/eval.py -> eval_test (batched, this repository), eval_validate (the reference's
EvalDataValidate, reached through the extended package path; it needs `open3d`, see dropin/open3d)."""
import _bootstrap  # noqa: F401
from regnet_for_3d_grasping_b200.grasp_eval import eval_test  # noqa: F401


def eval_validate(formal_dict, predicted_grasp, view_num, table_height, depth, width, gpu):
    from .eval_utils.evaluation_data_generator import EvalDataValidate
    view_cloud = EvalDataValidate(formal_dict, predicted_grasp, view_num, table_height, depth, width, gpu)
    return view_cloud.run_collision()
