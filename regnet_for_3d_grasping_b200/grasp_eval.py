"""View-collision filter of predicted grasps: `eval_test` of the reference
(dataset_utils/eval_score/eval.py:4-12 -> eval_utils/evaluation_data_generator.py:47-225, EvalDataTest.run_collision_view).

The reference walks the grasps in a Python loop (up to 4 000 per cloud, four times per cloud in utils.py:395-401); every
iteration transforms the whole view cloud into the gripper frame and counts the points in three regions.  Here the
grasps are processed in batches: one (batch, 3, N) transform and a handful of boolean reductions per batch, on whatever
device the inputs live on.  The normals the reference estimates in the constructor (open3d, every call) are never used by
this filter and are not computed.

Gripper constants: dataset_utils/eval_score/configs/config.py:11,27-41."""
import torch

NUM_POINTS_THRESHOLD = 16
CLOSE_REGION_MIN_POINTS = 16
NEIGHBOR_DEPTH = 0.005
BACK_COLLISION_THRESHOLD = 0.0
BACK_COLLISION_MARGIN = 0.0
FINGER_COLLISION_THRESHOLD = 0
FINGER_WIDTH = 0.01
HALF_HAND_THICKNESS = 0.005
BOTTOM_LENGTH = 0.06


def grasp_frames(grasp):
    """inv_transform_predicted_grasp (:115-163): grasp (B,8) = (centre, closing axis, angle, score) ->
    frame (B,3,3) with COLUMNS approach / closing axis / minor normal, centre (B,3), score (B,1)."""
    g = grasp.reshape(-1, 8).float()
    B = g.shape[0]
    center = g[:, :3].contiguous()
    angle = g[:, 6]
    cos_t, sin_t = torch.cos(angle), torch.sin(angle)
    zero, one = torch.zeros_like(cos_t), torch.ones_like(cos_t)
    R1 = torch.stack([cos_t, zero, -sin_t, zero, one, zero, sin_t, zero, cos_t], dim=1).view(B, 3, 3)

    def unit(v, fallback):
        n = torch.norm(v, dim=1)
        out = v / n.view(-1, 1)
        bad = n == 0
        if bad.any():
            out[bad] = torch.tensor(fallback, dtype=v.dtype, device=v.device)
        return out

    axis_y = unit(g[:, 3:6], [0.0, 1.0, 0.0])
    axis_x = unit(torch.stack([axis_y[:, 1], -axis_y[:, 0], zero], dim=1), [1.0, 0.0, 0.0])
    axis_z = unit(torch.cross(axis_x, axis_y, dim=1), [0.0, 0.0, 1.0])
    matrix = torch.bmm(torch.stack([axis_x, axis_y, axis_z], dim=2), R1)
    approach = unit(matrix[:, :, 0], [1.0, 0.0, 0.0])
    minor = torch.cross(approach, axis_y, dim=1)
    return torch.stack([approach, axis_y, minor], dim=2).contiguous(), center, g[:, 7].view(-1, 1).contiguous()


def view_collision_free(points, grasp, table_height, depth, width, batch=256):
    """Boolean (B,) mask of the grasps that pass finger_hand_view (:178-225): fingertips above the table, at least 16
    points between the hand's back plane and the fingertips, nothing behind the hand within its footprint, nothing inside
    the two finger volumes."""
    pts = points.float()
    frame, center, _ = grasp_frames(grasp)
    B = frame.shape[0]
    ok = ~(center[:, 2] + frame[:, 2, 0] * depth < table_height + 0.005)
    to_local = frame.transpose(1, 2).contiguous()
    shift = -torch.bmm(to_local, center.unsqueeze(2))                       # (B,3,1), as the reference builds it (:92-93)
    cloud = pts.t().contiguous()                                             # (3,N)
    half_w, half_s = width / 2 + FINGER_WIDTH, width / 2
    for lo in range(0, B, batch):
        hi = min(B, lo + batch)
        local = torch.matmul(to_local[lo:hi], cloud) + shift[lo:hi]         # (b,3,N)
        x, y, z = local[:, 0], local[:, 1], local[:, 2]
        close = (x > -BOTTOM_LENGTH) & (x < depth)
        in_z = (z < HALF_HAND_THICKNESS) & (z > -HALF_HAND_THICKNESS)
        back = close & (y < half_w) & (y > -half_w) & (x < -BACK_COLLISION_MARGIN) & in_z
        finger = close & in_z & (((y < half_w) & (y > half_s)) | ((y > -half_w) & (y < -half_s)))
        ok[lo:hi] &= (close.sum(1) >= NUM_POINTS_THRESHOLD) & ~(back.sum(1) > BACK_COLLISION_THRESHOLD) \
            & ~(finger.sum(1) > FINGER_COLLISION_THRESHOLD)
    return ok


def _region_tests(cloud, to_local, shift, depth, width):
    """The gripper-frame region tests shared by the view and the scene filters (:420-470, :480-525) for a batch of
    grasps: (local y (b,N), close-plane count, back-collision count, finger-collision count, closing-region mask)."""
    local = torch.matmul(to_local, cloud) + shift                            # (b,3,N)
    x, y, z = local[:, 0], local[:, 1], local[:, 2]
    d = depth.view(-1, 1) if isinstance(depth, torch.Tensor) else depth
    close = (x > -BOTTOM_LENGTH) & (x < d)
    half_w, half_s = width / 2 + FINGER_WIDTH, width / 2
    in_z = close & (z < HALF_HAND_THICKNESS) & (z > -HALF_HAND_THICKNESS)
    back = in_z & (y < half_w) & (y > -half_w) & (x < -BACK_COLLISION_MARGIN)
    finger = in_z & (((y < half_w) & (y > half_s)) | ((y > -half_w) & (y < -half_s)))
    region = in_z & (y < half_s) & (y > -half_s)
    return y, close.sum(1), back.sum(1), finger.sum(1), region


NORMAL_RADIUS, NORMAL_MAX_NN = 0.01, 30      # dataset_utils/eval_score/eval_utils/config.py


def _estimate_scene_normals(points):
    """pointcloud.py:27-43 of the reference's evaluation utilities."""
    import numpy as np
    import open3d
    cloud = open3d.geometry.PointCloud()
    cloud.points = open3d.utility.Vector3dVector(np.asarray(points, dtype=np.float64))
    cloud.estimate_normals(search_param=open3d.geometry.KDTreeSearchParamHybrid(radius=NORMAL_RADIUS, max_nn=NORMAL_MAX_NN),
                           fast_normal_computation=False)
    cloud.normalize_normals()
    cloud.orient_normals_towards_camera_location(np.zeros(3))
    return np.asarray(cloud.normals)


def eval_validate(formal_dict, predicted_grasp, view_num, table_height, depth, width, gpu, batch=128):
    """Drop-in for dataset_utils.eval_score.eval.eval_validate (EvalDataValidate.run_collision,
    evaluation_data_generator.py:231-538): grasps (B,8) -> (number of grasps without a SCENE collision, sum of their
    antipodal scores, number without a VIEW collision, those grasps, and the scene-collision-free ones).

    View filter = eval_test's with the validate variant's constants (fingertips may dip 5 mm below the table top, and the
    closing region must hold 16 view points).  Scene filter = the same region tests on the 8x denser scene cloud; a
    surviving grasp scores mean|n_y| over the outer `min(span / 3, 5 mm)` strip on the left of its closing region times
    the same on the right, n = scene normals in the gripper frame (:395-416).  `depth` may be a float or one value per
    grasp.  The view-cloud normals the reference estimates in its constructor are not used by anything and not computed."""
    import numpy as np
    as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
    grasp = as_t(predicted_grasp).float()
    dev = grasp.device
    if gpu != -1 and torch.cuda.is_available() and dev.type != "cuda":
        dev = torch.device("cuda", gpu)
    grasp = grasp.to(dev).view(-1, 8)
    view = as_t(formal_dict["view_cloud"]).float().to(dev)[:, :3].t().contiguous()
    scene = as_t(formal_dict["scene_cloud"]).float().to(dev)[:, :3].t().contiguous()
    if "scene_normal" in formal_dict:
        normals_np = formal_dict["scene_normal"]
    else:
        # the reference estimates them on the fly in this case (eval_utils/torch_scene_point_cloud.py:17-19 ->
        # pointcloud.py:27-43: hybrid kNN/radius search, normalised, oriented towards the origin): same recipe through
        # `open3d` -- the real package or this repository's stand-in -- cached on the scene dictionary for the next call
        normals_np = formal_dict.get("_estimated_scene_normal")
        if normals_np is None:
            normals_np = _estimate_scene_normals(np.asarray(as_t(formal_dict["scene_cloud"]).cpu())[:, :3])
            try:
                formal_dict["_estimated_scene_normal"] = normals_np
            except TypeError:      # an NpzFile is read-only: estimate again next time
                pass
    normal = as_t(normals_np).float().to(dev)[:, :3].t().contiguous()
    frame, center, _ = grasp_frames(grasp)
    B = frame.shape[0]
    per_grasp_depth = isinstance(depth, torch.Tensor) and depth.numel() > 1
    dep = depth.float().to(dev).view(-1) if per_grasp_depth else float(depth)
    to_local = frame.transpose(1, 2).contiguous()
    shift = -torch.bmm(to_local, center.unsqueeze(2))
    ok = ~(center[:, 2] + frame[:, 2, 0] * dep < table_height - 0.005)
    for lo in range(0, B, batch):
        hi = min(B, lo + batch)
        _, n_close, n_back, n_finger, region = _region_tests(view, to_local[lo:hi], shift[lo:hi],
                                                             dep[lo:hi] if per_grasp_depth else dep, width)
        ok[lo:hi] &= (n_close >= NUM_POINTS_THRESHOLD) & ~(n_back > BACK_COLLISION_THRESHOLD) \
            & ~(n_finger > FINGER_COLLISION_THRESHOLD) & (region.sum(1) >= CLOSE_REGION_MIN_POINTS)
    keep = torch.nonzero(ok).view(-1)
    grasp_view = grasp[keep]
    V = len(keep)
    free = torch.zeros(V, dtype=torch.bool, device=dev)
    score = torch.zeros(V, dtype=torch.float32, device=dev)
    inf = float("inf")
    for lo in range(0, V, batch):
        sel = keep[lo:lo + batch]
        y, n_close, n_back, n_finger, region = _region_tests(scene, to_local[sel], shift[sel],
                                                             dep[sel] if per_grasp_depth else dep, width)
        good = (n_close >= NUM_POINTS_THRESHOLD) & ~(n_back > BACK_COLLISION_THRESHOLD) \
            & ~(n_finger > FINGER_COLLISION_THRESHOLD) & (region.sum(1) >= CLOSE_REGION_MIN_POINTS)
        left_y = torch.where(region, y, torch.full_like(y, -inf)).amax(1, keepdim=True)
        right_y = torch.where(region, y, torch.full_like(y, inf)).amin(1, keepdim=True)
        strip = torch.minimum((left_y - right_y) / 3, torch.full_like(left_y, NEIGHBOR_DEPTH))
        ny = torch.matmul(to_local[sel][:, 1:2, :], normal)[:, 0].abs()      # |y component of the normals in the gripper frame|
        left = region & (y > left_y - strip)
        right = region & (y < right_y + strip)
        mean = lambda m: (ny * m).sum(1) / m.sum(1).clamp(min=1)
        free[lo:lo + len(sel)] = good
        score[lo:lo + len(sel)] = torch.where(good, mean(left) * mean(right), torch.zeros_like(mean(left)))
    return int(free.sum()), float(score.sum()), V, grasp_view, grasp_view[free]


def eval_test(points, predicted_grasp, view_num, table_height, depth, width, gpu):
    """Drop-in for dataset_utils.eval_score.eval.eval_test: points (N,3), predicted_grasp (B,8) -> the grasps without a
    view collision, in their original order.  `gpu` as in the reference (-1 = CPU, else that CUDA device); tensors that
    already live on a device are used where they are."""
    if isinstance(predicted_grasp, torch.Tensor):
        grasp = predicted_grasp.float()
    else:
        grasp = torch.as_tensor(predicted_grasp, dtype=torch.float32)
    dev = grasp.device
    if gpu != -1 and torch.cuda.is_available() and dev.type != "cuda":
        dev = torch.device("cuda", gpu)
    grasp = grasp.to(dev)
    pts = (points if isinstance(points, torch.Tensor) else torch.as_tensor(points)).to(dev)
    if grasp.numel() == 0:
        return grasp.view(-1, 8)
    return grasp.view(-1, 8)[view_collision_free(pts[:, :3], grasp, table_height, depth, width)]
