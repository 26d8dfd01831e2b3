"""Drop-in for the reference's nerf/volume_rendering_utils.py (alpha compositing + the Dex-NeRF
sigma-threshold depth), running as one CUDA kernel (dex-nerf_b200/csrc/composite.cu)."""
import torch

from . import _lib as L


_thr_cache = {}


def _thresholds_tensor(m_thres_cand, device):
    """Threshold candidates as a device tensor.  Cached per (values, device): building it from a Python
    list is a pageable host->device copy, i.e. a stream synchronisation in every render call."""
    if m_thres_cand is None:
        # the reference iterates over it unconditionally (volume_rendering_utils.py:53)
        raise TypeError("'NoneType' object is not iterable")
    vals = tuple(float(m) for m in m_thres_cand)
    if not vals:
        return None, 0
    key = (vals, str(device))
    t = _thr_cache.get(key)
    if t is None:
        if len(_thr_cache) > 64:
            _thr_cache.clear()
        t = torch.tensor(vals, dtype=torch.float32, device=device)
        _thr_cache[key] = t
    return t, len(vals)


def render_maps(radiance_field, depth_values, ray_directions, noise, white_background, thr, T,
                want_weights=True, want_indices=False):
    """Single launch of dexnerf_volume_render on flat (n,S) inputs; returns a dict of outputs."""
    rf = L.dev_f32(radiance_field, "radiance_field")
    z = L.dev_f32(depth_values, "depth_values")
    rd = L.dev_f32(ray_directions, "ray_directions")
    if rf.dim() != 3 or rf.shape[-1] != 4 or z.dim() != 2:
        # the reference's Dex gather depth_values[arange(n), idx] only works for 2-D depths (:57-58)
        raise IndexError("volume_render_radiance_field expects radiance_field (n,S,4) and depth_values (n,S)")
    n, S = z.shape
    if tuple(rf.shape[:2]) != (n, S) or tuple(rd.shape) != (n, 3):
        raise ValueError("shape mismatch: rf %s z %s rd %s" % (tuple(rf.shape), tuple(z.shape), tuple(rd.shape)))
    noise = L.dev_f32(noise, "noise", allow_none=True)
    dev = rf.device
    o = dict(rgb=torch.empty((n, 3), dtype=torch.float32, device=dev),
             disp=torch.empty((n,), dtype=torch.float32, device=dev),
             acc=torch.empty((n,), dtype=torch.float32, device=dev),
             depth=torch.empty((n,), dtype=torch.float32, device=dev),
             weights=torch.empty((n, S), dtype=torch.float32, device=dev) if want_weights else None,
             dex=torch.empty((T, n), dtype=torch.float32, device=dev) if T else None,
             dex_index=torch.empty((T, n), dtype=torch.int64, device=dev) if (T and want_indices) else None)
    if n:
        # algorithmic HBM bytes of this launch (24*S + 36 + 4*T per ray in the full form)
        nbytes = n * (S * (16 + 4 + (4 if noise is not None else 0) + (4 if want_weights else 0)) + 12 + 24 + 4 * T
                      + (8 * T if want_indices else 0))
        with L.timed("composite", n, S, nbytes):
            L.check(L.lib().dexnerf_volume_render(
                L.ptr(rf), L.ptr(z), L.ptr(rd), L.ptr(noise), n, S, int(bool(white_background)), L.ptr(thr), T,
                L.ptr(o["rgb"]), L.ptr(o["disp"]), L.ptr(o["acc"]), L.ptr(o["weights"]), L.ptr(o["depth"]),
                L.ptr(o["dex"]), L.ptr(o["dex_index"]), L.stream_ptr()), "volume_render_radiance_field")
    return o


def volume_render_radiance_field(radiance_field, depth_values, ray_directions,
                                 radiance_field_noise_std=0.0, white_background=False,
                                 m_thres_cand=None, noise=None):
    """volume_rendering_utils.py:6-70.  Returns
    (rgb_map, disp_map, acc_map, weights, depth_map, *depth_map_dex) with one Dex depth per
    threshold: the depth of the FIRST sample whose (noisy, rectified) sigma exceeds m, z[:, 0]
    when none does.  `noise` (extension) replays an already scaled N(0, std) draw instead of the
    internal torch.randn (:32-39)."""
    thr, T = _thresholds_tensor(m_thres_cand, radiance_field.device)
    if noise is None and radiance_field_noise_std > 0.0:
        noise = torch.randn(radiance_field.shape[:-1], dtype=radiance_field.dtype,
                            device=radiance_field.device) * radiance_field_noise_std
    o = render_maps(radiance_field, depth_values, ray_directions, noise, white_background, thr, T)
    dex = [o["dex"][t] for t in range(T)]
    return tuple([o["rgb"], o["disp"], o["acc"], o["weights"], o["depth"]] + dex)
