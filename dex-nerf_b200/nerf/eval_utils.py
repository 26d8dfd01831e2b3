"""The callers on the far side of the hot path:

* the evaluation render loop of eval_nerf.py:166-206 over a camera path
  (`pose_spherical`, load_blender.py:33-38), updated to the fork's 5-argument `get_ray_bundle` and
  (6 + T)-tuple API, sharded over GPUs by image rows;
* the validation-time depth metrics of train_dexnerf_rgb.py:391-428 / train_utils.py:9-30, computed
  on the device for all T threshold planes at once instead of T host round trips.
"""
import math

import torch

from . import _lib as L


# ------------------------------------------------------------------ camera path
def pose_spherical(theta, phi, radius):
    """load_blender.py:33-38: camera-to-world pose on a sphere of `radius` looking at the origin
    (OpenGL camera convention).  float32 (4, 4) CPU tensor."""
    def trans_t(t):
        m = torch.eye(4, dtype=torch.float32)
        m[2, 3] = t
        return m

    def rot_phi(p):
        m = torch.eye(4, dtype=torch.float32)
        m[1, 1] = m[2, 2] = math.cos(p)
        m[1, 2] = -math.sin(p)
        m[2, 1] = -m[1, 2]
        return m

    def rot_theta(th):
        m = torch.eye(4, dtype=torch.float32)
        m[0, 0] = m[2, 2] = math.cos(th)
        m[0, 2] = -math.sin(th)
        m[2, 0] = -m[0, 2]
        return m

    c2w = trans_t(radius)
    c2w = rot_phi(phi / 180.0 * math.pi) @ c2w
    c2w = rot_theta(theta / 180.0 * math.pi) @ c2w
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ c2w


def world2cam_from_blender_pose(c2w):
    """The fork's get_ray_bundle takes an OpenCV world->cam extrinsic (nerf_helpers.py:67-112) while
    load_blender.py produces OpenGL cam->world poses: flip the camera's y and z axes, then invert."""
    c2w = torch.as_tensor(c2w, dtype=torch.float64)
    if c2w.shape[0] == 3:
        c2w = torch.cat((c2w, torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)), 0)
    cv = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], dtype=torch.float64))
    return torch.linalg.inv(cv).to(torch.float32).contiguous()      # (LAPACK hands back column-major strides)


def render_poses_spherical(n_frames=40, phi=-30.0, radius=4.0):
    """load_blender.py:116-121: the 360 degree path the reference's eval renders."""
    step = 360.0 / n_frames
    return torch.stack([pose_spherical(-180.0 + i * step, phi, radius) for i in range(n_frames)], 0)


def render_path(poses_w2c, height, width, intrinsic, model_coarse, model_fine, options, encode_position_fn,
                encode_direction_fn, m_thres_cand=(), rank=0, world=1, gather=False, on_frame=None):
    """eval_nerf.py:166-206 as a function: render every pose of `poses_w2c` (P, 4, 4 world->cam) in
    validation mode under no_grad.  With world > 1 each rank renders its block of image rows
    (nerf.row_block, no collective); gather=True all-gathers the finished planes.
    Returns (frames, seconds_per_frame): frames[i] is the (6 + T)-tuple of run_one_iter_of_nerf for
    this rank's rows (or the full frame when gathered / world == 1); seconds_per_frame is measured
    with CUDA events around each frame, like the reference's per-image timer."""
    from .nerf_helpers import get_ray_bundle
    from .sharding import gather_rows, row_block
    from .train_utils import run_one_iter_of_nerf
    K_host = torch.as_tensor(intrinsic, dtype=torch.float32).cpu()
    focal = float(K_host[0, 0])
    K = K_host.cuda()
    row0, rows = row_block(height, rank, world)
    thr = list(m_thres_cand)
    frames, times = [], []
    # one upload for the whole path: a pageable host->device copy per frame would synchronise the stream
    # and expose the launch latency of every frame's first kernels
    poses_dev = torch.stack([torch.as_tensor(p, dtype=torch.float32) for p in poses_w2c], 0).cuda()
    for i in range(poses_dev.shape[0]):
        pose = poses_dev[i]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        with torch.no_grad():
            ro, rd = get_ray_bundle(height, width, None, pose, K, row_start=row0, row_count=rows)
            out = run_one_iter_of_nerf(height, width, focal, model_coarse, model_fine, ro, rd, options,
                                       mode="validation", encode_position_fn=encode_position_fn,
                                       encode_direction_fn=encode_direction_fn, m_thres_cand=thr)
            if gather and world > 1:
                out = tuple(gather_rows(o, height) if o is not None else None for o in out)
        e1.record()
        frames.append(out)
        times.append((e0, e1))
        if on_frame is not None:
            on_frame(i, out)
    torch.cuda.synchronize()
    secs = [a.elapsed_time(b) * 1e-3 for a, b in times]
    return frames, secs


def cast_to_image(tensor):
    """eval_nerf.py:19-31: (H, W, 3) float image -> uint8 HWC numpy array, clipped."""
    img = (tensor.detach().clamp(0.0, 1.0) * 255.0).to(torch.uint8)
    return img.cpu().numpy()


# ------------------------------------------------------------------ validation depth metrics
def dex_depth_error_metrics(depth_planes, depth_gt, mask=None):
    """All threshold planes at once.  depth_planes: (T, ...) tensor or a sequence of T tensors (the
    tail of run_one_iter_of_nerf's tuple), depth_gt: (...), mask: bool/uint8 (...) or None for the
    reference's (gt > 0) & (gt < 1.25) (train_dexnerf_rgb.py:392).
    Returns (metrics (T, 4) CUDA tensor [abs err mm, err2, err4, err8], best index CUDA int32)."""
    if not isinstance(depth_planes, torch.Tensor):
        depth_planes = torch.stack(list(depth_planes), 0)
    T = depth_planes.shape[0]
    pred = L.dev_f32(_to_device(depth_planes.reshape(T, -1)), "depth_planes")
    gt = L.dev_f32(_to_device(depth_gt.reshape(-1)), "depth_gt")
    n = gt.numel()
    if pred.shape[1] != n:
        raise ValueError("depth planes have %d pixels, ground truth %d" % (pred.shape[1], n))
    m = None
    if mask is not None:
        m = mask.reshape(-1).to(device=gt.device, dtype=torch.uint8).contiguous()
        if m.numel() != n:
            raise ValueError("mask has %d pixels, ground truth %d" % (m.numel(), n))
    out = torch.empty((T, 4), dtype=torch.float32, device=gt.device)
    best = torch.empty((), dtype=torch.int32, device=gt.device)
    ws = torch.empty(4 * T + 1, dtype=torch.float64, device=gt.device)
    L.check(L.lib().dexnerf_depth_error_metrics(L.ptr(pred), L.ptr(gt), L.ptr(m), n, T, L.ptr(out), L.ptr(best),
                                                L.ptr(ws), L.stream_ptr()), "depth_error_metrics")
    L.launch_count += 1      # two kernels behind the call
    return out, best


def _to_device(t):
    """The reference's validation block hands these helpers `.cpu()` tensors (train_dexnerf_rgb.py:391-415:
    `depth_target.cpu()`, `...detach().cpu()`); the metrics run on the device, so host tensors are uploaded
    to the current CUDA device here (the inputs are one depth map each - this is reporting, not the hot path)."""
    t = t.detach()
    if not t.is_cuda:
        t = t.to(device=torch.device("cuda", torch.cuda.current_device()))
    return t.to(torch.float32)


def compute_err_metric(depth_gt, depth_pred, mask):
    """train_utils.py:9-30, same signature and result dict (python floats).  CPU tensors are accepted, as the
    reference script passes them (train_dexnerf_rgb.py:391-404)."""
    out, _ = dex_depth_error_metrics(depth_pred.reshape(1, -1), depth_gt, mask)
    v = out[0].tolist()
    return {"depth_abs_err": v[0], "depth_err2": v[1], "depth_err4": v[2], "depth_err8": v[3]}


def depth_error_img(D_est_tensor, D_gt_tensor, mask, abs_thres=1.0, dilate_radius=1):
    """train_utils.py:45-70: the colour-coded depth error image of the validation block
    (train_dexnerf_rgb.py:415), same signature and return value - a numpy (H, W, 3) float32 image of batch
    entry 0 with the colour legend in the top-left corner - computed by one kernel on the device
    (csrc/metrics.cu); `dilate_radius` is unused, as in the reference."""
    if D_est_tensor.dim() != 3 or D_gt_tensor.shape != D_est_tensor.shape or mask.shape != D_est_tensor.shape:
        raise ValueError("depth_error_img expects (B, H, W) depth maps and a (B, H, W) mask")
    est = L.dev_f32(_to_device(D_est_tensor[0]), "D_est_tensor")
    gt = L.dev_f32(_to_device(D_gt_tensor[0]), "D_gt_tensor")
    m = mask.detach()[0].to(device=est.device, dtype=torch.uint8).contiguous()
    H, W = est.shape
    out = torch.empty((H, W, 3), dtype=torch.float32, device=est.device)
    L.check(L.lib().dexnerf_depth_error_image(L.ptr(est), L.ptr(gt), L.ptr(m), H, W, float(abs_thres), L.ptr(out),
                                              L.stream_ptr()), "depth_error_img")
    return out.cpu().numpy()


def select_dex_threshold(depth_planes, depth_gt, mask=None):
    """The selection loop of train_dexnerf_rgb.py:393-404: the first threshold whose mean absolute
    depth error (mm) is the smallest and below 1000.  Returns (index, err dict, metrics (T, 4))
    with ONE device->host read of 4T + 1 numbers."""
    out, best = dex_depth_error_metrics(depth_planes, depth_gt, mask)
    host = torch.cat((out.reshape(-1), best.to(torch.float32).reshape(1))).cpu()
    idx = int(host[-1])
    table = host[:-1].reshape(-1, 4)
    err = None
    if idx >= 0:
        v = table[idx].tolist()
        err = {"depth_abs_err": v[0], "depth_err2": v[1], "depth_err4": v[2], "depth_err8": v[3]}
    return idx, err, table
