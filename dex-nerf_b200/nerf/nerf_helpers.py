"""Drop-in for the reference's nerf/nerf_helpers.py: same names, signatures and results, with the
arithmetic done by the sm_100a kernels behind include/dexnerf.h.  Citations are to
/root/reference/nerf-pytorch/nerf/nerf_helpers.py unless stated."""
import math
from typing import Optional

import torch

from . import _lib as L


def img2mse(img_src, img_tgt):
    """nerf_helpers.py:9-10."""
    return torch.nn.functional.mse_loss(img_src, img_tgt)


def mse2psnr(mse):
    """nerf_helpers.py:13-17 (a zero loss is replaced by 1e-5)."""
    if mse == 0:
        mse = 1e-5
    return -10.0 * math.log10(mse)


def get_minibatches(inputs: torch.Tensor, chunksize: Optional[int] = 1024 * 8):
    """nerf_helpers.py:20-25: list of row chunks."""
    return [inputs[i:i + chunksize] for i in range(0, inputs.shape[0], chunksize)]


def meshgrid_xy(tensor1: torch.Tensor, tensor2: torch.Tensor) -> (torch.Tensor, torch.Tensor):
    """np.meshgrid(..., indexing="xy") (nerf_helpers.py:28-40): ii[r, c] = tensor1[c],
    jj[r, c] = tensor2[r].  Pure index plumbing - views, no kernel."""
    ii = tensor1[None, :].expand(tensor2.shape[0], tensor1.shape[0])
    jj = tensor2[:, None].expand(tensor2.shape[0], tensor1.shape[0])
    return ii, jj


def _small_to_device(t, name, shape):
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t, dtype=torch.float32)
    if tuple(t.shape) != shape:
        raise ValueError("%s must have shape %s, got %s" % (name, shape, tuple(t.shape)))
    return t.to(device="cuda", dtype=torch.float32).contiguous()


def cumprod_exclusive(tensor: torch.Tensor) -> torch.Tensor:
    """tf.math.cumprod(..., exclusive=True) along the last dim (nerf_helpers.py:43-64), as a
    warp-level exclusive product scan."""
    x = L.dev_f32(tensor, "tensor")
    S = x.shape[-1]
    out = torch.empty_like(x)
    if x.numel():
        L.check(L.lib().dexnerf_cumprod_exclusive(L.ptr(x), x.numel() // S, S, L.ptr(out), L.stream_ptr()),
                "cumprod_exclusive")
    return out


def get_ray_bundle(height: int, width: int, focal_length, tform_cam2world: torch.Tensor,
                   intrinsic: torch.Tensor, row_start: int = 0, row_count: Optional[int] = None):
    """nerf_helpers.py:67-112.  Same quirks as the reference: `focal_length` is ignored, the 4x4
    is a world->cam extrinsic that is inverted here, both pixel axes are divided by
    intrinsic[0, 0], and the result is (height, width, 3).  `row_start`/`row_count` (extension)
    produce only a block of image rows - what each GPU of a row-sharded render asks for."""
    T = _small_to_device(tform_cam2world, "tform_cam2world", (4, 4))
    K = _small_to_device(intrinsic, "intrinsic", (3, 3))
    rows = height - row_start if row_count is None else row_count
    ro = torch.empty((rows, width, 3), dtype=torch.float32, device="cuda")
    rd = torch.empty_like(ro)
    L.check(L.lib().dexnerf_ray_bundle(L.ptr(T), L.ptr(K), height, width, row_start, rows, L.ptr(ro),
                                       L.ptr(rd), L.stream_ptr()), "get_ray_bundle")
    return ro, rd


class _Embedder:
    """What get_embedding_function returns: callable like the reference's lambda
    (nerf_helpers.py:162-169) but with its three settings readable, so run_one_iter_of_nerf can fuse
    the encoding into the MLP kernel instead of materialising it."""

    def __init__(self, num_encoding_functions, include_input, log_sampling):
        self.num_encoding_functions = int(num_encoding_functions)
        self.include_input = bool(include_input)
        self.log_sampling = bool(log_sampling)

    @property
    def out_dim(self):
        return (3 if self.include_input else 0) + 6 * self.num_encoding_functions

    def __call__(self, x):
        return positional_encoding(x, self.num_encoding_functions, self.include_input, self.log_sampling)


def positional_encoding(tensor, num_encoding_functions=6, include_input=True, log_sampling=True) -> torch.Tensor:
    """nerf_helpers.py:115-159: [x, sin(f0 x), cos(f0 x), sin(f1 x), ...] on the last dim (size 3)."""
    x = L.dev_f32(tensor, "tensor")
    if x.shape[-1] != 3:
        raise ValueError("positional_encoding expects (..., 3) points, got %s" % (tuple(x.shape),))
    if num_encoding_functions == 0:
        if include_input:
            return tensor  # the reference returns the input itself (:156-157)
        raise RuntimeError("positional_encoding: nothing to encode")
    D = (3 if include_input else 0) + 6 * num_encoding_functions
    out = torch.empty(x.shape[:-1] + (D,), dtype=torch.float32, device=x.device)
    if x.numel():
        L.check(L.lib().dexnerf_positional_encoding(L.ptr(x), x.numel() // 3, int(num_encoding_functions),
                                                    int(bool(include_input)), int(bool(log_sampling)),
                                                    L.ptr(out), L.stream_ptr()), "positional_encoding")
    return out


def get_embedding_function(num_encoding_functions=6, include_input=True, log_sampling=True):
    """nerf_helpers.py:162-169."""
    return _Embedder(num_encoding_functions, include_input, log_sampling)


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """nerf_helpers.py:172-199."""
    ro, rd = L.dev_f32(rays_o, "rays_o"), L.dev_f32(rays_d, "rays_d")
    oo, od = torch.empty_like(ro), torch.empty_like(rd)
    if ro.numel():
        L.check(L.lib().dexnerf_ndc_rays(L.ptr(ro), L.ptr(rd), ro.numel() // 3, int(H), int(W), float(focal),
                                         float(near), L.ptr(oo), L.ptr(od), L.stream_ptr()), "ndc_rays")
    return oo, od


def _sample_pdf_impl(bins, weights, num_samples, det, u=None, return_indices=False):
    b, w = L.dev_f32(bins, "bins"), L.dev_f32(weights, "weights")
    if b.dim() != 2 or w.dim() != 2 or w.shape[1] != b.shape[1] - 1 or w.shape[0] != b.shape[0]:
        raise ValueError("sample_pdf expects bins (n, B) and weights (n, B-1); got %s and %s"
                         % (tuple(b.shape), tuple(w.shape)))
    n, B = b.shape
    if u is None and not det:
        u = torch.rand((n, num_samples), dtype=torch.float32, device=b.device)
    u = L.dev_f32(u, "u", allow_none=True)
    if u is not None and tuple(u.shape) != (n, num_samples):
        raise ValueError("u must have shape (n, num_samples)")
    samples = torch.empty((n, num_samples), dtype=torch.float32, device=b.device)
    inds = torch.empty((n, num_samples), dtype=torch.int64, device=b.device) if return_indices else None
    L.check(L.lib().dexnerf_sample_pdf(L.ptr(b), L.ptr(w), n, B, int(num_samples), L.ptr(u), L.ptr(samples),
                                       L.ptr(inds), L.stream_ptr()), "sample_pdf")
    return (samples, inds) if return_indices else samples


def sample_pdf_2(bins, weights, num_samples, det=False, u=None, return_indices=False):
    """nerf_helpers.py:262-304 - this is what `nerf.sample_pdf` resolves to in the reference
    (train_utils.py:6 re-binds the name).  The external torchsearchsorted call (:290) is part of
    the kernel.  Extensions: `u` replays a given uniform draw, `return_indices` also returns the
    searchsorted indices (int64)."""
    return _sample_pdf_impl(bins, weights, num_samples, det, u, return_indices)


def sample_pdf(bins, weights, num_samples, det=False, u=None, return_indices=False):
    """nerf_helpers.py:224-259 (the gather_cdf_util variant): bit-identical to sample_pdf_2 in the
    reference, so both names reach the same kernel."""
    return _sample_pdf_impl(bins, weights, num_samples, det, u, return_indices)


def gather_cdf_util(cdf, inds):
    """nerf_helpers.py:202-221: row-wise gather with out-of-range indices clamped and zeroed.
    Kept for API completeness (index plumbing only; sample_pdf no longer needs it)."""
    valid = inds < cdf.shape[1]
    clamped = torch.where(valid, inds, torch.full_like(inds, cdf.shape[1] - 1))
    flat = torch.gather(cdf, 1, clamped.reshape(inds.shape[0], -1)).reshape(inds.shape)
    return flat * valid.to(flat.dtype)
