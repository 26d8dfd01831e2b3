"""Ray-cache files in the reference's on-disk format.

cache_dataset.py:104-135 writes, per training image, `torch.save({"height", "width", "focal_length",
"ray_bundle" (2, ..., 3) = stack(origins, directions), "target"})` and, per validation image,
`{"height", "width", "focal_length", "ray_origins", "ray_directions", "target"}`; the training loop
reads them back at train_dexnerf_rgb.py:186-204 and draws `num_random_rays` of the cached rays without
replacement.  Here the bundles come from this repo's ray-generation kernel with the fork's 5-argument
`get_ray_bundle` (the reference's cache script still calls the old 4-argument form and cannot run)."""
import numpy as np
import torch

from .nerf_helpers import get_ray_bundle


def train_cache_entry(height, width, focal_length, pose_w2c, intrinsic, image, num_random_rays=None, rng=None):
    """cache_dataset.py:69-115.  `num_random_rays=None` is the script's --sample-all; otherwise that many
    pixels are drawn without replacement (np.random.choice like the script, or `rng`)."""
    ro, rd = get_ray_bundle(height, width, focal_length, pose_w2c, intrinsic)
    target = torch.as_tensor(image)
    if num_random_rays is not None:
        rng = rng or np.random
        sel = torch.from_numpy(rng.choice(height * width, size=(num_random_rays,), replace=False))
        rows, cols = (sel // width).to(ro.device), (sel % width).to(ro.device)
        ro, rd = ro[rows, cols, :], rd[rows, cols, :]
        target = target.to(ro.device)[rows, cols, :]
    return {"height": height, "width": width, "focal_length": focal_length,
            "ray_bundle": torch.stack([ro, rd], dim=0).detach().cpu(), "target": target.detach().cpu()}


def val_cache_entry(height, width, focal_length, pose_w2c, intrinsic, image):
    """cache_dataset.py:121-135."""
    ro, rd = get_ray_bundle(height, width, focal_length, pose_w2c, intrinsic)
    return {"height": height, "width": width, "focal_length": focal_length, "ray_origins": ro.detach().cpu(),
            "ray_directions": rd.detach().cpu(), "target": torch.as_tensor(image).detach().cpu()}


def save_cache_entry(entry, path):
    torch.save(entry, path)


def load_cache_entry(path):
    return torch.load(path, weights_only=False)


def training_rays_from_cache(cache_dict, num_random_rays, device="cuda", rng=None):
    """train_dexnerf_rgb.py:186-204: (ray_origins, ray_directions, target) of `num_random_rays` cached rays
    drawn without replacement, on `device`."""
    bundle = cache_dict["ray_bundle"].to(device)
    ro, rd = bundle[0].reshape((-1, 3)), bundle[1].reshape((-1, 3))
    target = cache_dict["target"][..., :3].reshape((-1, 3))
    rng = rng or np.random
    sel = torch.from_numpy(rng.choice(ro.shape[0], size=(num_random_rays,), replace=False))
    return ro[sel.to(device)].contiguous(), rd[sel.to(device)].contiguous(), target[sel].to(device).contiguous()
