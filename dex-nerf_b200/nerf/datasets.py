"""Dataset readers of the reference's `nerf` namespace: the three names the
training scripts import next to the render functions (train_dexnerf_rgb.py:15-19, :61-76):

    load_blender_data      nerf/load_blender.py:41-127     NeRF-synthetic: transforms_<split>.json + PNGs
    load_messytable_data   nerf/load_messytable.py:17-176  Dex-NeRF scenes: <split>/<scene>/{meta.pkl, image, depth}
    load_llff_data         nerf/load_llff.py:266-354       forward-facing captures: poses_bounds.npy + images/

Host-side I/O, as in the reference (it is Python there too): they return CPU tensors / numpy arrays with the
reference's shapes, dtypes and quirks, which the training loop then moves to the GPU.  Images are decoded with
PIL (the reference's `imageio.imread` decodes PNG / JPEG to the same pixels); resizing is cv2 INTER_AREA /
INTER_NEAREST like the reference.  Quirks kept on purpose, each pinned by tests/test_datasets.py against the
reference's own output on byte-identical synthetic datasets:

* blender `half_res` divides H, W and the focal length by FOUR (load_blender.py:105-109); RGBA stays RGBA;
  `camera_angle_x` is read from the last split; `render_poses` are float64 (an int matrix times float32).
* messytable: scenes are visited in `os.listdir` order; grey-scale images become three equal channels;
  depth is millimetres / 1000; H, W are always halved and the focal length (taken from the LAST scene read)
  divided by four (:143-146) even when `half_res=False`; `half_res` quarters the intrinsics and pins the
  principal point to (240, 135) (:66-71).
* llff: poses are re-ordered [y, -x, z], rescaled by 1 / (min bound * bd_factor), recentred through float64
  and stored back as float32; the hold-out view is the pose closest to the average pose.

Two differences, both where the reference cannot run: a missing `images_<factor>` directory is produced
in-process with cv2 INTER_AREA (the reference shells out to ImageMagick's `mogrify`, :27-66), and
`path_zflat=True` uses an integer view count (the reference passes 60.0 to np.linspace, which raises)."""
import json
import os
import pickle

import cv2
import numpy as np
import torch
from PIL import Image

from .eval_utils import pose_spherical

_SPLITS = ("train", "val", "test")


def _imread(path, mode=None):
    img = Image.open(path)
    if mode is not None:
        img = img.convert(mode)
    return np.array(img)


def _render_poses():
    """The 40-view turntable every loader returns (load_blender.py:82-88): float64, like the reference's
    int-matrix @ float32 product."""
    return torch.stack([torch.from_numpy(np.asarray(pose_spherical(a, -30.0, 4.0), dtype=np.float64))
                        for a in np.linspace(-180, 180, 40 + 1)[:-1]], 0)


def _resize_all(arrays, W, H, interpolation):
    return torch.stack([torch.from_numpy(cv2.resize(a, dsize=(W, H), interpolation=interpolation)) for a in arrays], 0)


def _split_indices(counts):
    edges = np.concatenate([[0], np.cumsum(counts)])
    return [np.arange(edges[i], edges[i + 1]) for i in range(len(counts))]


def load_blender_data(basedir, half_res=False, testskip=1, debug=False):
    """load_blender.py:41-127 -> (imgs (N,H,W,4) float32, poses (N,4,4) float32, render_poses (40,4,4),
    [H, W, focal], i_split)."""
    imgs, poses, counts, meta = [], [], [], None
    for s in _SPLITS:
        with open(os.path.join(basedir, "transforms_%s.json" % s), "r") as fp:
            meta = json.load(fp)
        step = 1 if (s == "train" or testskip == 0) else testskip
        frames = meta["frames"][::step]
        imgs.append((np.array([_imread(os.path.join(basedir, f["file_path"] + ".png")) for f in frames]) / 255.0)
                    .astype(np.float32))
        poses.append(np.array([np.array(f["transform_matrix"]) for f in frames]).astype(np.float32))
        counts.append(len(frames))
    imgs, poses = np.concatenate(imgs, 0), np.concatenate(poses, 0)
    H, W = imgs[0].shape[:2]
    focal = 0.5 * W / np.tan(0.5 * float(meta["camera_angle_x"]))
    if debug:
        H, W, focal = H // 32, W // 32, focal / 32.0
        out = _resize_all(imgs, 25, 25, cv2.INTER_AREA)          # the reference's fixed 25x25 thumbnails
    else:
        if half_res:
            H, W, focal = H // 4, W // 4, focal / 4.0
        out = _resize_all(imgs, W, H, cv2.INTER_AREA)
    return out, torch.from_numpy(poses), _render_poses(), [H, W, focal], _split_indices(counts)


def load_messytable_data(basedir, half_res=False, testskip=1, debug=False, imgname="0128_irL_kuafu_half.png",
                         is_real_rgb=False):
    """load_messytable.py:17-176 -> (imgs (N,H,W,3), poses (N,4,4) world->cam extrinsics, render_poses,
    [H, W, focal], i_split, intrinsics (N,3,3), depths (N,H,W) in metres)."""
    depth_name, ext_key, int_key = (("depth.png", "extrinsic", "intrinsic") if is_real_rgb
                                    else ("depthL.png", "extrinsic_l", "intrinsic_l"))
    imgs, poses, intrinsics, depths, counts, meta = [], [], [], [], [], None
    for s in _SPLITS:
        split_dir = os.path.join(basedir, s)
        n = 0
        for scene in os.listdir(split_dir):
            with open(os.path.join(split_dir, scene, "meta.pkl"), "rb") as fp:
                meta = pickle.load(fp)
            img = _imread(os.path.join(split_dir, scene, imgname))
            if img.ndim != 3:
                img = np.repeat(img[..., None], 3, axis=-1)
            imgs.append(img)
            depths.append(_imread(os.path.join(split_dir, scene, depth_name)) / 1000)
            poses.append(np.array(meta[ext_key]))
            K = np.array(meta[int_key])
            if half_res:
                K[:2, :] = K[:2, :] / 4
                K[0, 2], K[1, 2] = 240.0, 135.0
            intrinsics.append(K)
            n += 1
        counts.append(n)
    imgs = (np.array(imgs) / 255.0).astype(np.float32)
    poses = np.array(poses).astype(np.float32)
    intrinsics = np.array(intrinsics).astype(np.float32)
    depths = np.array(depths).astype(np.float32)
    H, W = imgs[0].shape[:2]
    focal = meta[int_key][0, 0]
    if debug:
        H, W, focal = H // 32, W // 32, focal / 32.0
        out_i, out_d = _resize_all(imgs, 25, 25, cv2.INTER_AREA), _resize_all(depths, 25, 25, cv2.INTER_NEAREST)
    else:
        H, W, focal = H // 2, W // 2, focal / 4.0
        out_i, out_d = _resize_all(imgs, W, H, cv2.INTER_AREA), _resize_all(depths, W, H, cv2.INTER_NEAREST)
    return (out_i, torch.from_numpy(poses), _render_poses(), [H, W, focal], _split_indices(counts),
            torch.from_numpy(intrinsics), out_d)


# ------------------------------------------------------------------------------------------ LLFF
_IMG_EXT = ("JPG", "jpg", "png")


def _image_files(d):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(_IMG_EXT)]


def _minify(basedir, factor):
    """images_<factor>/ from images/ when it is missing (see the module docstring)."""
    dst = os.path.join(basedir, "images_%d" % factor)
    if os.path.exists(dst):
        return
    os.makedirs(dst)
    for f in _image_files(os.path.join(basedir, "images")):
        img = _imread(f, "RGB")
        small = cv2.resize(img, dsize=(img.shape[1] // factor, img.shape[0] // factor), interpolation=cv2.INTER_AREA)
        Image.fromarray(small).save(os.path.join(dst, os.path.splitext(os.path.basename(f))[0] + ".png"))


def _unit(v):
    return v / np.linalg.norm(v)


def _look_at(z, up, pos):
    """3x4 camera-to-world with the given viewing axis, approximate up vector and position."""
    z = _unit(z)
    x = _unit(np.cross(up, z))
    y = _unit(np.cross(z, x))
    return np.stack([x, y, z, pos], 1)


def _average_pose(poses):
    """3x5 [R | t | hwf] of the mean camera (load_llff.py:157-166)."""
    # the viewing axis is normalised here AND inside _look_at, as in the reference (the second division by a
    # norm of 1 +- 1 ulp is not always the identity)
    return np.concatenate([_look_at(_unit(poses[:, :3, 2].sum(0)), poses[:, :3, 1].sum(0), poses[:, :3, 3].mean(0)),
                           poses[0, :3, -1:]], 1)


def _homogeneous(p34):
    row = np.tile(np.reshape([0, 0, 0, 1.0], [1, 1, 4]), [p34.shape[0], 1, 1])
    return np.concatenate([p34, row], 1)


def _recenter(poses):
    """Express every pose in the frame of the average pose (load_llff.py:187-201; the 4x4 work is float64)."""
    out = poses + 0
    avg = np.concatenate([_average_pose(poses)[:3, :4], np.reshape([0, 0, 0, 1.0], [1, 4])], -2)
    out[:, :3, :4] = (np.linalg.inv(avg) @ _homogeneous(poses[:, :3, :4]))[:, :3, :4]
    return out


def _spiral_path(c2w, up, rads, focal, zrate, rots, n_views):
    """load_llff.py:169-184: cameras on a spiral around the average pose, all looking at depth `focal`."""
    rads = np.array(list(rads) + [1.0])
    hwf = c2w[:, 4:5]
    look = np.dot(c2w[:3, :4], np.array([0, 0, -focal, 1.0]))
    out = []
    for theta in np.linspace(0.0, 2.0 * np.pi * rots, n_views + 1)[:-1]:
        c = np.dot(c2w[:3, :4], np.array([np.cos(theta), -np.sin(theta), -np.sin(theta * zrate), 1.0]) * rads)
        out.append(np.concatenate([_look_at(_unit(c - look), up, c), hwf], 1))
    return out


def _spherify(poses, bds):
    """load_llff.py:204-263: recentre on the point closest to all optical axes, scale the rig to the unit
    sphere and return a 120-view circle at the cameras' mean height."""
    d, o = poses[:, :3, 2:3], poses[:, :3, 3:4]
    A = np.eye(3) - d * np.transpose(d, [0, 2, 1])
    b = -A @ o
    center = np.squeeze(-np.linalg.inv((np.transpose(A, [0, 2, 1]) @ A).mean(0)) @ b.mean(0))
    up = (poses[:, :3, 3] - center).mean(0)
    v0 = _unit(up)
    v1 = _unit(np.cross([0.1, 0.2, 0.3], v0))
    v2 = _unit(np.cross(v0, v1))
    frame = np.stack([v1, v2, v0, center], 1)
    reset = np.linalg.inv(_homogeneous(frame[None])) @ _homogeneous(poses[:, :3, :4])
    rad = np.sqrt(np.mean(np.sum(np.square(reset[:, :3, 3]), -1)))
    sc = 1.0 / rad
    reset[:, :3, 3] *= sc
    bds *= sc
    rad *= sc
    zh = np.mean(reset[:, :3, 3], 0)[2]
    radcircle = np.sqrt(rad ** 2 - zh ** 2)
    circle = []
    for th in np.linspace(0.0, 2.0 * np.pi, 120):
        cam = np.array([radcircle * np.cos(th), radcircle * np.sin(th), zh])
        z = _unit(cam)
        x = _unit(np.cross(z, np.array([0, 0, -1.0])))
        y = _unit(np.cross(z, x))
        circle.append(np.stack([x, y, z, cam], 1))
    circle = np.stack(circle, 0)
    hwf = poses[0, :3, -1:]
    circle = np.concatenate([circle, np.broadcast_to(hwf, circle[:, :3, -1:].shape)], -1)
    reset = np.concatenate([reset[:, :3, :4], np.broadcast_to(hwf, reset[:, :3, -1:].shape)], -1)
    return reset, circle, bds


def load_llff_data(basedir, factor=8, recenter=True, bd_factor=0.75, spherify=False, path_zflat=False):
    """load_llff.py:266-354 -> (images (N,H,W,3) float32, poses (N,3,5) float32 [R | t | hwf], bds (N,2),
    render_poses float32, i_test)."""
    arr = np.load(os.path.join(basedir, "poses_bounds.npy"))
    poses = arr[:, :-2].reshape([-1, 3, 5]).transpose([1, 2, 0])
    bds = arr[:, -2:].transpose([1, 0])
    suffix = ""
    if factor is not None:
        suffix = "_%d" % factor
        _minify(basedir, factor)
    else:
        factor = 1
    files = _image_files(os.path.join(basedir, "images" + suffix))
    if poses.shape[-1] != len(files):
        raise ValueError("load_llff_data: %d images but %d poses in %s" % (len(files), poses.shape[-1], basedir))
    shape = _imread(files[0], "RGB").shape
    poses[:2, 4, :] = np.array(shape[:2]).reshape([2, 1])
    poses[2, 4, :] = poses[2, 4, :] * 1.0 / factor
    imgs = np.stack([_imread(f, "RGB")[..., :3] / 255.0 for f in files], -1)

    poses = np.concatenate([poses[:, 1:2, :], -poses[:, 0:1, :], poses[:, 2:, :]], 1)
    poses = np.moveaxis(poses, -1, 0).astype(np.float32)
    images = np.moveaxis(imgs, -1, 0).astype(np.float32)
    bds = np.moveaxis(bds, -1, 0).astype(np.float32)
    sc = 1.0 if bd_factor is None else 1.0 / (bds.min() * bd_factor)
    poses[:, :3, 3] *= sc
    bds *= sc
    if recenter:
        poses = _recenter(poses)
    if spherify:
        poses, render_poses, bds = _spherify(poses, bds)
    else:
        c2w = _average_pose(poses)
        up = _unit(poses[:, :3, 1].sum(0))
        close_depth, inf_depth = bds.min() * 0.9, bds.max() * 5.0
        dt = 0.75
        focal = 1.0 / ((1.0 - dt) / close_depth + dt / inf_depth)
        rads = np.percentile(np.abs(poses[:, :3, 3]), 90, 0)
        n_views, n_rots = 120, 2
        if path_zflat:
            c2w[:3, 3] = c2w[:3, 3] + (-close_depth * 0.1) * c2w[:3, 2]
            rads[2] = 0.0
            n_views, n_rots = 60, 1
        render_poses = _spiral_path(c2w, up, rads, focal, 0.5, n_rots, n_views)
    render_poses = np.array(render_poses).astype(np.float32)
    c2w = _average_pose(poses)
    i_test = np.argmin(np.sum(np.square(c2w[:3, 3] - poses[:, :3, 3]), -1))
    return images.astype(np.float32), poses.astype(np.float32), bds, render_poses, i_test
