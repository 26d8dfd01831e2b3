"""Tiny synthetic datasets in the three on-disk formats the reference's loaders read
(nerf/load_blender.py, nerf/load_messytable.py, nerf/load_llff.py), written deterministically
(seeded numpy, lossless PNG) so that tests/golden/make_golden.py (reference loaders) and
tests/test_datasets.py (this repo's loaders) see byte-identical files."""
import json
import os
import pickle

import numpy as np
from PIL import Image


def _rigid(rs):
    q, _ = np.linalg.qr(rs.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    m = np.eye(4)
    m[:3, :3] = q
    m[:3, 3] = rs.randn(3) * 2
    return m


def build_blender(root, seed=1):
    """transforms_{train,val,test}.json + RGBA PNGs (load_blender.py:41-63); 32x32 images, 3/2/4 frames."""
    rs = np.random.RandomState(seed)
    os.makedirs(root, exist_ok=True)
    for split, n in (("train", 3), ("val", 2), ("test", 4)):
        os.makedirs(os.path.join(root, split), exist_ok=True)
        frames = []
        for i in range(n):
            img = rs.randint(0, 256, size=(32, 32, 4), dtype=np.uint8)
            Image.fromarray(img, "RGBA").save(os.path.join(root, split, "r_%d.png" % i))
            frames.append({"file_path": "./%s/r_%d" % (split, i), "transform_matrix": _rigid(rs).tolist()})
        with open(os.path.join(root, "transforms_%s.json" % split), "w") as fp:
            json.dump({"camera_angle_x": 0.6911112070083618, "frames": frames}, fp)
    return root


def build_messytable(root, seed=2, imgname="0128_irL_kuafu_half.png"):
    """{train,val,test}/<scene>/{meta.pkl, <imgname>, depthL.png, depth.png} (load_messytable.py:39-78):
    a grey-scale IR image (expanded to three channels by the loader), 16-bit depth in millimetres, 4x4
    extrinsics and 3x3 intrinsics in meta.pkl.  One scene per split: os.listdir order is then unambiguous."""
    rs = np.random.RandomState(seed)
    H, W = 36, 64
    for split in ("train", "val", "test"):
        d = os.path.join(root, split, "0-300002-%s" % split)
        os.makedirs(d, exist_ok=True)
        Image.fromarray(rs.randint(0, 256, size=(H, W), dtype=np.uint8), "L").save(os.path.join(d, imgname))
        Image.fromarray(rs.randint(0, 256, size=(H, W, 3), dtype=np.uint8), "RGB").save(os.path.join(d, "rgb.png"))
        for name in ("depthL.png", "depth.png"):
            depth = rs.randint(300, 4000, size=(H, W)).astype(np.uint16)
            Image.fromarray(depth).save(os.path.join(d, name))
        K = np.array([[1386.4, 0.0, 960.0], [0.0, 1386.4, 540.0], [0.0, 0.0, 1.0]]) + rs.rand(3, 3) * 1e-3
        meta = {"extrinsic_l": _rigid(rs), "intrinsic_l": K, "extrinsic": _rigid(rs), "intrinsic": K * 1.01}
        with open(os.path.join(d, "meta.pkl"), "wb") as fp:
            pickle.dump(meta, fp)
    return root


def build_llff(root, seed=3, n=6, factor=8):
    """poses_bounds.npy (n x 17: a 3x5 [R | t | hwf] block + near/far bounds), images/ and the already
    minified images_<factor>/ (load_llff.py:69-138; the reference shells out to `mogrify` when the
    minified directory is missing, which this fixture avoids)."""
    rs = np.random.RandomState(seed)
    H, W = 48, 64
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    os.makedirs(os.path.join(root, "images_%d" % factor), exist_ok=True)
    rows = []
    for i in range(n):
        Image.fromarray(rs.randint(0, 256, size=(H, W, 3), dtype=np.uint8), "RGB").save(
            os.path.join(root, "images", "img_%03d.png" % i))
        Image.fromarray(rs.randint(0, 256, size=(H // factor, W // factor, 3), dtype=np.uint8), "RGB").save(
            os.path.join(root, "images_%d" % factor, "img_%03d.png" % i))
        m = _rigid(rs)
        # forward-facing rig: cameras near the origin looking roughly down -z
        m[:3, :3] = np.eye(3) + 0.05 * rs.randn(3, 3)
        m[:3, 3] = 0.3 * rs.randn(3)
        block = np.concatenate([m[:3, :4], np.array([[H], [W], [55.0]])], axis=1)   # 3 x 5
        rows.append(np.concatenate([block.reshape(-1), np.array([1.2 + 0.1 * rs.rand(), 9.0 + rs.rand()])]))
    np.save(os.path.join(root, "poses_bounds.npy"), np.stack(rows, 0))
    return root
