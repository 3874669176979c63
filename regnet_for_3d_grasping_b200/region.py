"""Region stage of REGNet on device: grasp-centre selection and region cropping,
plus the grouped max-pool (R3) and the per-row masked sampler of the closing-box crop (R6).

`get_grasp_allobj` keeps the signature and return tuple of dataset_utils/get_regiondataset.py:13-42 so that
train.py:237 / test.py:135 can call it unchanged; the implementation is three batched kernel launches instead of
B*N_C Python iterations with host synchronisation (csrc/region.cu, C ABI in include/regnet_b200.h section 3).

Random draws come from a counter-based device generator seeded from torch's CPU generator (`torch.manual_seed`
makes them reproducible); the reference draws from numpy's global state seeded with the wall clock
(train.py:59, test.py:56), so only the distributions -- not the streams -- can be matched.
"""
import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _seed(seed):
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return ctypes.c_uint64(seed & (2 ** 64 - 1))


def _need(t, name, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (the region stage has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}")


def select_score_center(pc, pre_score, center_num, score_thre, seed=None, return_count=False):
    """get_regiondataset.py:354-434.  pc (B,N,6), pre_score (B,N) -> center_pc (B,center_num,6),
    center_pc_index (B,center_num) int64."""
    _need(pc, "pc")
    _need(pre_score, "pre_score")
    B, N, C = pc.shape
    if C != 6:
        pc = pc[:, :, :6]
    pc = pc.contiguous()
    score = pre_score.contiguous().view(B, N)
    lib = _lib.load()
    center_index = torch.empty(B, center_num, dtype=torch.int64, device=pc.device)
    center_pc = torch.empty(B, center_num, 6, dtype=torch.float32, device=pc.device)
    count = torch.empty(B, dtype=torch.int32, device=pc.device)
    ws_bytes = lib.regnet_select_score_center_workspace(B, N, center_num)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pc.device)
    with torch.cuda.device(pc.device):
        _lib.check(lib.regnet_select_score_center(_p(pc), _p(score), B, N, int(center_num), float(score_thre), _seed(seed),
                                                  _p(center_index), _p(center_pc), _p(count), _p(ws), ws_bytes,
                                                  _lib.current_stream_ptr()))
    if return_count:
        return center_pc, center_index, count
    return center_pc, center_index


def get_group_pc(pc, center_pc, center_pc_index, group_num, width, height, depth, r_time, seed=None, return_count=False):
    """get_regiondataset.py:311-352.  -> pc_group_index (B,N_C,group_num) int64, pc_group (B,N_C,group_num,6)."""
    _need(pc, "pc")
    _need(center_pc, "center_pc")
    B, N, C = pc.shape
    if C != 6:
        pc = pc[:, :, :6]
    pc = pc.contiguous()
    center_pc = center_pc.contiguous()
    NC = center_pc.shape[1]
    radius = max(width, height, depth) * r_time            # python double, rounded to fp32 at the comparison
    index = torch.empty(B, NC, group_num, dtype=torch.int64, device=pc.device)
    group = torch.empty(B, NC, group_num, 6, dtype=torch.float32, device=pc.device)
    count = torch.empty(B, NC, dtype=torch.int32, device=pc.device)
    with torch.cuda.device(pc.device):
        lib = _lib.load()
        nbytes = int(lib.regnet_ball_crop_workspace_bytes(B, N))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=pc.device)      # scratch of the uniform grid
        _lib.check(lib.regnet_ball_crop_sample_ws(_p(pc), _p(center_pc), B, N, NC, float(radius), int(group_num),
                                                  _seed(seed), _p(index), _p(group), _p(count), _p(ws), nbytes,
                                                  _lib.current_stream_ptr()))
    if return_count:
        return index, group, count
    return index, group


_SCENE_CACHE = {}          # (path, mtime, size, device) -> annotation tensors; insertion-ordered, oldest evicted
_SCENE_CACHE_MAX = 2048
_SCENE_CACHE_MAX_BYTES = int(float(__import__("os").environ.get("REGNET_SCENE_CACHE_MB", "512")) * 2 ** 20)
_scene_cache_bytes = 0


def _load_scene_grasps(path, device):
    """The grasp annotations of one scene file: (frames (G,4,4), score, antipodal, centre score) as fp32 tensors on
    `device`, for both on-disk formats of the reference (get_regiondataset.py:63-87).  The reference un-pickles the file
    and uploads it inside every training step (0.6 ms per scene); here the tensors of the last 2 048 scenes stay
    resident, keyed by path + mtime + size (REGNET_NO_SCENE_CACHE=1 disables it)."""
    import os
    st = os.stat(path)
    key = (os.path.abspath(path), st.st_mtime_ns, st.st_size, str(device))
    use_cache = os.environ.get("REGNET_NO_SCENE_CACHE", "0") != "1"
    if use_cache and key in _SCENE_CACHE:
        return _SCENE_CACHE[key]
    out = _read_scene_grasps(path, device)
    if use_cache:
        # bounded by entries AND by bytes (512 MB of device memory by default, REGNET_SCENE_CACHE_MB): scenes with long
        # grasp lists must not pin gigabytes next to a training run
        global _scene_cache_bytes
        nbytes = sum(t.numel() * t.element_size() for t in set(out))
        while _SCENE_CACHE and (len(_SCENE_CACHE) >= _SCENE_CACHE_MAX or _scene_cache_bytes + nbytes > _SCENE_CACHE_MAX_BYTES):
            old = _SCENE_CACHE.pop(next(iter(_SCENE_CACHE)))
            _scene_cache_bytes -= sum(t.numel() * t.element_size() for t in set(old))
        if nbytes <= _SCENE_CACHE_MAX_BYTES:
            _SCENE_CACHE[key] = out
            _scene_cache_bytes += nbytes
    return out


def _read_scene_grasps(path, device):
    import numpy as np
    data = np.load(path, allow_pickle=True)
    t = lambda v: (torch.as_tensor(np.asarray(v), dtype=torch.float32) if not isinstance(v, torch.Tensor) else v.float()).to(device)
    if "frame" in data.keys():
        score = t(data["antipodal_score"])
        return t(data["frame"]), score, score, score
    return (t(data["select_frame"]), t(data["select_antipodal_score"]), t(data["select_antipodal_score"]),
            t(data["select_center_score"]))


def transform_grasp(grasp_ori, grasp_score_ori, antipodal_score_ori, center_score_ori):
    """get_regiondataset.py:136-199: (B,N_C,3,4) frames [x | y | z | centre] -> (B,N_C,10) = (centre, closing axis with
    non-negative x, angle in (-pi, pi], score, antipodal score, centre score); rows without a grasp stay -1 (8 columns
    when no scene carries antipodal scores, like the reference)."""
    import math
    B, CN = grasp_score_ori.shape
    cols = 8 if bool((antipodal_score_ori == -1).all()) else 10
    out = torch.full((B, CN, cols), -1.0, device=grasp_ori.device)
    axis_x = grasp_ori[:, :, :3, 0].reshape(B * CN, 3)
    axis_y = grasp_ori[:, :, :3, 1].reshape(B * CN, 3).clone()
    axis_z = grasp_ori[:, :, :3, 2].reshape(B * CN, 3)
    no_grasp = (axis_x == -1).all(dim=1)
    angle = torch.atan2(axis_x[:, 2], axis_z[:, 2])
    flip = axis_y[:, 0] < 0
    angle = torch.where(flip, math.pi - angle, angle)
    axis_y = torch.where(flip[:, None], -axis_y, axis_y)
    angle = torch.where(angle >= 2 * math.pi, angle - 2 * math.pi, angle)
    angle = torch.where(angle <= -2 * math.pi, angle + 2 * math.pi, angle)
    angle = torch.where(angle > math.pi, angle - 2 * math.pi, angle)
    angle = torch.where(angle <= -math.pi, angle + 2 * math.pi, angle)
    angle = torch.where(no_grasp, torch.full_like(angle, -1.0), angle)
    out[:, :, :3] = grasp_ori[:, :, :3, 3]
    out[:, :, 3:6] = axis_y.view(B, CN, 3)
    out[:, :, 6] = angle.view(B, CN)
    out[:, :, 7] = grasp_score_ori
    if cols > 8:
        out[:, :, 8] = antipodal_score_ori
        out[:, :, 9] = center_score_ori
    return out


def get_center_grasp(center_pc_index, center_pc, data_paths, depth, use_theta=True):
    """get_regiondataset.py:45-134 (_get_center_grasp): the annotated grasp nearest to every centre (squared distance by
    the reference's |a|^2 + |b|^2 - 2ab expansion, kept iff it does not exceed 0.005) -> grasp_labels (B,N_C,10)
    (or (B,N_C,13) frames + score with use_theta=False); centres without a grasp are -1.  One scene file per cloud; all clouds
    are handled by one batched (N_C x G) distance matrix on the centres' device."""
    from torch.nn.utils.rnn import pad_sequence
    dev = center_pc.device
    B, NC = center_pc_index.shape
    label = torch.full((B, NC, 3, 4), -1.0, device=dev)
    score_l = torch.full((B, NC), -1.0, device=dev)
    anti_l = torch.full((B, NC), -1.0, device=dev)
    cent_l = torch.full((B, NC), -1.0, device=dev)
    n = len(data_paths)
    if n > 0:
        # all scenes at once: annotations padded to the longest list, padded slots pushed far away
        loaded = [_load_scene_grasps(path, dev) for path in data_paths]
        counts = torch.tensor([len(x[0]) for x in loaded], device=dev)
        grasp = pad_sequence([x[0][:, :3, :4] for x in loaded], batch_first=True)            # (n, G, 3, 4)
        score, anti, cent = (pad_sequence([x[k] for x in loaded], batch_first=True) for k in (1, 2, 3))
        G = grasp.shape[1]
        gx = grasp[:, :, :, 0]
        centre = (grasp[:, :, :, 3] + gx * depth).float()
        centre = (centre - gx * depth).float()               # the reference's round trip (:92-93)
        pad = torch.arange(G, device=dev).view(1, G) >= counts.view(n, 1)
        centre = torch.where(pad[:, :, None], torch.full_like(centre, 1.0e6), centre)
        p1 = center_pc[:n, :, :3].float()
        d = -2 * torch.bmm(p1, centre.transpose(2, 1))       # |a|^2 + |b|^2 - 2ab, in the reference's order (:268-272)
        d = d + (centre * centre).sum(2).view(n, 1, G)
        d = d + (p1 * p1).sum(2).view(n, NC, 1)
        dmin, arg = torch.min(d.double(), dim=2)
        keep = ~(dmin > 0.005)
        take = lambda t: t.gather(1, arg)
        label[:n] = torch.where(keep[:, :, None, None], grasp.gather(1, arg[:, :, None, None].expand(n, NC, 3, 4)), label[:n])
        score_l[:n] = torch.where(keep, take(score), score_l[:n])
        anti_l[:n] = torch.where(keep, take(anti), anti_l[:n])
        cent_l[:n] = torch.where(keep, take(cent), cent_l[:n])
    if use_theta:
        return transform_grasp(label, score_l, anti_l, cent_l)
    flat = label.view(-1, 3, 4)
    inv = flat[:, 0, 1] < 0
    flat[inv, :, 1:2] = -flat[inv, :, 1:2]
    out = torch.full((B, NC, 13), -1.0, device=dev)
    out[:, :, :12] = flat.transpose(2, 1).contiguous().view(B, NC, 12)
    out[:, :, 12:13] = score_l.view(B, NC, 1)
    return out


def get_grasp_allobj(pc, predict_score, params, data_paths, use_theta=True, seed=None):
    """Drop-in for dataset_utils.get_regiondataset.get_grasp_allobj (get_regiondataset.py:13-42).

    Returns (center_pc, center_pc_index, pc_group_index, pc_group, pc_group_more_index, pc_group_more, grasp_labels);
    grasp_labels is None for data_paths=[] (the inference path of test.py:135), else the label lookup of
    get_center_grasp over the scenes' annotation files (one path per cloud)."""
    (center_num, score_thre, group_num, r_time_group, group_num_more, r_time_group_more, width, height, depth) = params
    s = None if seed is None else int(seed)
    center_pc, center_pc_index = select_score_center(pc, predict_score, center_num, score_thre, seed=s)
    pc_group_index, pc_group = get_group_pc(pc, center_pc, center_pc_index, group_num, width, height, depth, r_time_group,
                                            seed=None if s is None else s + 1)
    pc_group_more_index, pc_group_more = get_group_pc(pc, center_pc, center_pc_index, group_num_more, width, height, depth,
                                                      r_time_group_more, seed=None if s is None else s + 2)
    grasp_labels = None
    if len(data_paths) > 0:
        grasp_labels = get_center_grasp(center_pc_index, center_pc, data_paths, depth, use_theta)
    return center_pc, center_pc_index, pc_group_index, pc_group, pc_group_more_index, pc_group_more, grasp_labels


def sample_mask_rows(mask, num, min_count=5, seed=None, return_count=False):
    """gripper_region_network.py:532-544: for each row of a (rows,G) boolean mask pick `num` column indices:
    more than `num` set -> without replacement; more than `min_count` -> with replacement; else -1 (row rejected)."""
    if not mask.is_cuda:
        raise RuntimeError("mask must be a CUDA tensor")
    m = mask.to(torch.uint8).contiguous()
    rows, G = m.shape
    index = torch.empty(rows, num, dtype=torch.int64, device=m.device)
    count = torch.empty(rows, dtype=torch.int32, device=m.device)
    if rows == 0:                      # an empty tensor has a null data pointer: nothing to ask the library
        return (index, count) if return_count else index
    with torch.cuda.device(m.device):
        _lib.check(_lib.load().regnet_mask_sample(_p(m), rows, G, int(num), int(min_count), _seed(seed), _p(index), _p(count),
                                                  _lib.current_stream_ptr()))
    if return_count:
        return index, count
    return index


def closing_box_mask(group_points, centre, rot, x_limit, y_limit, z_limit):
    """(M,G) uint8 membership of the gripper's closing box (gripper_region_network.py:505-528) in one pass:
    group_points (M,G,C>=3) fp32, centre (M,3), rot (M,3,3); x_limit / y_limit: python floats or (M,) / (M,1) tensors."""
    _need(group_points, "group_points")
    pts = group_points.contiguous()
    M, G, C = pts.shape
    centre = centre.float().contiguous()
    rot = rot.float().contiguous()
    xr = x_limit.float().reshape(-1).contiguous() if isinstance(x_limit, torch.Tensor) else None
    yr = y_limit.float().reshape(-1).contiguous() if isinstance(y_limit, torch.Tensor) else None
    if (xr is not None and xr.numel() != M) or (yr is not None and yr.numel() != M):
        raise RuntimeError("closing_box_mask: per-grasp limits must have one value per grasp")
    mask = torch.empty(M, G, dtype=torch.uint8, device=pts.device)
    if M == 0 or G == 0:
        return mask
    with torch.cuda.device(pts.device):
        for lo in range(0, M, 65535):
            hi = min(M, lo + 65535)
            _lib.check(_lib.load().regnet_closing_box_mask(
                _p(pts[lo:hi]), hi - lo, G, C, _p(centre[lo:hi]), _p(rot[lo:hi]), _p(xr[lo:hi]) if xr is not None else None,
                _p(yr[lo:hi]) if yr is not None else None, 0.0 if xr is not None else float(x_limit),
                0.0 if yr is not None else float(y_limit), float(z_limit), _p(mask[lo:hi]), _lib.current_stream_ptr()))
    return mask


def gather_max(all_feature, index):
    """max over each group's feature rows: all_feature (B,N,C) point-major contiguous, index (B,N_C,G) int64
    -> (B,N_C,C).  Equals MaxPool1d(G)(all_feature.view(-1,C)[index + b*N].permute(0,2,1)) of
    gripper_region_network.py:389-395 + utils/pointnet2.py:161,167."""
    _need(all_feature, "all_feature")
    _need(index, "index", torch.int64)
    f = all_feature.contiguous()
    B, N, C = f.shape
    ix = index.contiguous()
    _, NC, G = ix.shape
    out = torch.empty(B, NC, C, dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _lib.check(_lib.load().regnet_gather_max(_p(f), _p(ix), B, N, NC, G, C, _p(out), _lib.current_stream_ptr()))
    return out
