"""CPU oracle for the Dex-NeRF ray-render hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU (torch-on-CPU, fp32) restatement of the reference's algorithm for the
path SURVEY.md section 8(a) lists.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it, and only as the checker or the
timed CPU baseline.  Nothing under `dex-nerf_b200/` imports it; the product path is CUDA-only.

Parity status: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: `tests/golden/make_golden.py`
imports the unmodified reference from /root/reference (with the three import shims in
`tests/golden/_shims`), runs it on seeded inputs and commits the results as `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks every function here against those files.
The one third-party dependency on the path, `torchsearchsorted` (unpinned git HEAD,
nerf-pytorch/requirements.txt:9), is absent from /root/reference; its published contract
(`side="right"`: first index i with cdf[i] > u) is restated in `searchsorted_right`.

Deliberate, documented differences from the reference's arithmetic (all at the last-ulp level):
  * reductions whose order the reference leaves to ATen (`sum`, `cumsum`, `cumprod`) are done
    here in float64, sequentially, and rounded to float32 once per output element.  CPU ATen
    does exactly that for cumsum/cumprod (acc_type<float> = double) and a vectorised cascade
    for `sum`; the CUDA kernels follow this file's definition so that indices are bit-exact.
  * the two 8-layer model forwards are the REPAIRED forwards of SURVEY.md section 8(a-3): the
    reference's `FlexibleNeRFModel.forward` reads a non-existent attribute
    (nerf/models.py:243) and `PaperNeRFModel.forward` feeds xyz+dir to the xyz trunk
    (nerf/models.py:165-169); both raise for the 8x256 configuration.

All citations are relative to /root/reference/nerf-pytorch/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

F32 = torch.float32
F64 = torch.float64


# --------------------------------------------------------------------------------------
# a-1  ray generation                                   nerf/nerf_helpers.py:28-40, 67-112
# --------------------------------------------------------------------------------------
def meshgrid_xy(xs: torch.Tensor, ys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """'xy'-indexed grid: ii[r, c] = xs[c], jj[r, c] = ys[r]  (nerf_helpers.py:28-40)."""
    ii = xs[None, :].expand(ys.shape[0], xs.shape[0])
    jj = ys[:, None].expand(ys.shape[0], xs.shape[0])
    return ii, jj


def get_ray_bundle(height: int, width: int, focal_length, tform_world2cam: torch.Tensor,
                   intrinsic: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """nerf_helpers.py:67-112.  `focal_length` is ignored; the 4x4 is world->cam and is inverted;
    BOTH pixel axes are divided by K[0,0] (fx) - the reference's quirk at :100-101.
    Returns (H, W, 3) origins and directions."""
    T = tform_world2cam.to(F32)
    K = intrinsic.to(F32)
    ii, jj = meshgrid_xy(torch.arange(width, dtype=F32), torch.arange(height, dtype=F32))
    dx = (ii - K[0, 2]) / K[0, 0]
    dy = (jj - K[1, 2]) / K[0, 0]
    dz = torch.ones_like(dx)
    Rinv = torch.inverse(T[:3, :3])
    # rd[a] = sum_b d[b] * Rinv[a, b], summed left to right in fp32
    rd = torch.stack(
        [(dx * Rinv[a, 0] + dy * Rinv[a, 1]) + dz * Rinv[a, 2] for a in range(3)], dim=-1)
    ro = torch.inverse(T)[:3, 3].expand(rd.shape)
    return ro.contiguous(), rd.contiguous()


def ndc_rays(H: int, W: int, focal: float, near: float, rays_o: torch.Tensor,
             rays_d: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Forward-facing NDC re-parameterisation (nerf_helpers.py:172-199)."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    o = rays_o + t[..., None] * rays_d
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    o0 = sx * o[..., 0] / o[..., 2]
    o1 = sy * o[..., 1] / o[..., 2]
    o2 = 1.0 + 2.0 * near / o[..., 2]
    d0 = sx * (rays_d[..., 0] / rays_d[..., 2] - o[..., 0] / o[..., 2])
    d1 = sy * (rays_d[..., 1] / rays_d[..., 2] - o[..., 1] / o[..., 2])
    d2 = -2.0 * near / o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


# --------------------------------------------------------------------------------------
# a-2  positional encoding                                 nerf/nerf_helpers.py:115-169
# --------------------------------------------------------------------------------------
def frequency_bands(num_encoding_functions: int, log_sampling: bool = True) -> torch.Tensor:
    L = num_encoding_functions
    if L == 0:
        return torch.zeros(0, dtype=F32)
    if log_sampling:
        return 2.0 ** torch.linspace(0.0, L - 1, L, dtype=F32)
    return torch.linspace(1.0, 2.0 ** (L - 1), L, dtype=F32)


def positional_encoding(x: torch.Tensor, num_encoding_functions: int = 6,
                        include_input: bool = True, log_sampling: bool = True) -> torch.Tensor:
    """Columns [x, sin(f0 x), cos(f0 x), sin(f1 x), cos(f1 x), ...] (nerf_helpers.py:115-159).
    With L == 0 and include_input the input itself is returned (:156-157)."""
    parts: List[torch.Tensor] = [x] if include_input else []
    for f in frequency_bands(num_encoding_functions, log_sampling):
        arg = x * f
        parts.append(torch.sin(arg))
        parts.append(torch.cos(arg))
    if len(parts) == 1:
        return parts[0]
    return torch.cat(parts, dim=-1)


# --------------------------------------------------------------------------------------
# a-4  stratified sampler                                     nerf/train_utils.py:104-136
# --------------------------------------------------------------------------------------
def stratified_z(near: torch.Tensor, far: torch.Tensor, num_coarse: int, lindisp: bool = False,
                 t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """near, far: (n, 1).  t_rand: optional (n, Nc) uniform draws (the reference's torch.rand at
    train_utils.py:132); None means perturb == False."""
    t = torch.linspace(0.0, 1.0, num_coarse, dtype=F32)
    if not lindisp:
        z = near * (1.0 - t) + far * t
    else:
        z = 1.0 / (1.0 / near * (1.0 - t) + 1.0 / far * t)
    z = z.expand(near.shape[0], num_coarse)
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat((mids, z[..., -1:]), dim=-1)
        lower = torch.cat((z[..., :1], mids), dim=-1)
        z = lower + (upper - lower) * t_rand
    return z.contiguous()


# --------------------------------------------------------------------------------------
# a-5 / a-6  compositing + Dex-NeRF threshold depth
#            nerf/nerf_helpers.py:43-64, nerf/volume_rendering_utils.py:6-70
# --------------------------------------------------------------------------------------
def cumprod_exclusive(x: torch.Tensor) -> torch.Tensor:
    """T_i = prod_{j<i} x_j, accumulated sequentially in fp64, rounded per element."""
    c = torch.cumprod(x.to(F64), dim=-1)
    out = torch.ones_like(c)
    out[..., 1:] = c[..., :-1]
    return out.to(F32)


def volume_render_radiance_field(radiance_field: torch.Tensor, depth_values: torch.Tensor,
                                 ray_directions: torch.Tensor,
                                 radiance_field_noise_std: float = 0.0,
                                 white_background: bool = False,
                                 m_thres_cand: Optional[Sequence[float]] = None,
                                 noise: Optional[torch.Tensor] = None):
    """volume_rendering_utils.py:6-70.  `noise` (n, S), if given, is the already scaled
    N(0, std) draw the reference makes internally at :32-39 (so tests can replay it).
    Returns (rgb_map, disp_map, acc_map, weights, depth_map, *dex_depths)."""
    if m_thres_cand is None:
        raise TypeError("m_thres_cand is required (the reference iterates over it, :53)")
    z = depth_values
    rd64 = ray_directions.to(F64)
    rd_norm = torch.sqrt((rd64 * rd64).sum(-1)).to(F32)
    dists = torch.cat((z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)), dim=-1)
    dists = dists * rd_norm[..., None]
    rgb = torch.sigmoid(radiance_field[..., :3])
    raw_sigma = radiance_field[..., 3]
    if noise is not None:
        raw_sigma = raw_sigma + noise
    elif radiance_field_noise_std > 0.0:
        raw_sigma = raw_sigma + torch.randn(raw_sigma.shape, dtype=F32) * radiance_field_noise_std
    sigma = torch.relu(raw_sigma)
    alpha = 1.0 - torch.exp(-sigma * dists)
    weights = alpha * cumprod_exclusive(1.0 - alpha + 1e-10)

    w64 = weights.to(F64)
    rgb_map = (w64[..., None] * rgb.to(F64)).sum(dim=-2).to(F32)
    depth_map = (w64 * z.to(F64)).sum(dim=-1).to(F32)
    acc_map = w64.sum(dim=-1).to(F32)
    disp_map = 1.0 / torch.max(torch.full_like(depth_map, 1e-10), depth_map / acc_map)
    if white_background:
        rgb_map = rgb_map + (1.0 - acc_map[..., None])

    dex = []
    rows = torch.arange(z.shape[0])
    for m in m_thres_cand:
        idx = dex_first_crossing(sigma, float(m))
        dex.append(z[rows, idx])
    return (rgb_map, disp_map, acc_map, weights, depth_map, *dex)


def dex_first_crossing(sigma: torch.Tensor, m: float) -> torch.Tensor:
    """Index of the FIRST sample with sigma > m (strict), 0 when none crosses
    (argmax of a 0/1 row, volume_rendering_utils.py:54-56).  int64."""
    hit = sigma > m
    first = torch.argmax(hit.to(torch.int32), dim=-1)
    return first.to(torch.int64)


# --------------------------------------------------------------------------------------
# a-7  hierarchical resampling                nerf/nerf_helpers.py:262-304 (sample_pdf_2)
# --------------------------------------------------------------------------------------
def searchsorted_right(cdf: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """torchsearchsorted.searchsorted(cdf, u, side='right'): first i with cdf[row, i] > u,
    len(cdf) when none.  Restated from the library's documented contract."""
    return (cdf[:, None, :] <= u[:, :, None]).sum(-1).to(torch.int64)


def pdf_to_cdf(weights: torch.Tensor) -> torch.Tensor:
    w = weights + 1e-5
    total = w.to(F64).sum(-1, keepdim=True).to(F32)
    pdf = w / total
    cdf = torch.cumsum(pdf.to(F64), dim=-1).to(F32)
    return torch.cat((torch.zeros_like(cdf[..., :1]), cdf), dim=-1)


def sample_pdf(bins: torch.Tensor, weights: torch.Tensor, num_samples: int, det: bool = False,
               u: Optional[torch.Tensor] = None, return_indices: bool = False):
    """`nerf.sample_pdf` == sample_pdf_2 (nerf_helpers.py:262-304; binding train_utils.py:6).
    bins (n, B), weights (n, B-1).  `u` (n, Nf) replays the reference's torch.rand draw."""
    cdf = pdf_to_cdf(weights)
    n = cdf.shape[0]
    if u is None:
        if det:
            u = torch.linspace(0.0, 1.0, num_samples, dtype=F32).expand(n, num_samples)
        else:
            u = torch.rand(n, num_samples, dtype=F32)
    u = u.contiguous()
    inds = searchsorted_right(cdf, u)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    samples = bin_b + t * (bin_a - bin_b)
    if return_indices:
        return samples, inds
    return samples


def merge_fine(z_coarse: torch.Tensor, z_samples: torch.Tensor) -> torch.Tensor:
    """train_utils.py:170-173: sort(cat(z_coarse, z_samples)) values."""
    return torch.sort(torch.cat((z_coarse, z_samples), dim=-1), dim=-1).values


# --------------------------------------------------------------------------------------
# a-3  models (functional, on a state_dict)                     nerf/models.py:123-256
# --------------------------------------------------------------------------------------
def _linear(x: torch.Tensor, sd: Dict[str, torch.Tensor], name: str, bf16: bool) -> torch.Tensor:
    w, b = sd[name + ".weight"].to(F32), sd[name + ".bias"].to(F32)
    if bf16:  # bf16 operands, fp32 accumulate: the tensor-core kernel's arithmetic contract
        x = x.to(torch.bfloat16).to(F32)
        w = w.to(torch.bfloat16).to(F32)
    return x @ w.t() + b


def flexible_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, skip_connect_every: int = 4,
                     use_viewdirs: bool = True, bf16: bool = False) -> torch.Tensor:
    """REPAIRED FlexibleNeRFModel.forward (models.py:233-256): skip condition is
    `i % skip == 0 and i > 0` (what __init__ builds, :210); cat order (x, xyz) (:245);
    layer1 has no ReLU (:238); alpha comes from the trunk output, not from feat (:248-249).
    With bf16=True every tensor-core layer sees bf16-rounded operands; fc_alpha and fc_rgb stay
    fp32 (they run on CUDA cores in the kernel)."""
    dim_xyz = sd["layer1.weight"].shape[1]
    xyz, view = x[..., :dim_xyz], x[..., dim_xyz:]
    n_xyz = len([k for k in sd if k.startswith("layers_xyz.") and k.endswith(".weight")])
    h = _linear(xyz, sd, "layer1", bf16)
    for i in range(n_xyz):
        if i % skip_connect_every == 0 and i > 0:
            h = torch.cat((h, xyz), dim=-1)
        h = torch.relu(_linear(h, sd, f"layers_xyz.{i}", bf16))
    if not use_viewdirs:
        return _linear(h, sd, "fc_out", False)     # a head: fp32 on the CUDA cores, like fc_alpha / fc_rgb
    feat = torch.relu(_linear(h, sd, "fc_feat", bf16))
    alpha = _linear(h, sd, "fc_alpha", False)
    y = torch.relu(_linear(torch.cat((feat, view), dim=-1), sd, "layers_dir.0", bf16))
    rgb = _linear(y, sd, "fc_rgb", False)
    return torch.cat((rgb, alpha), dim=-1)


def paper_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, use_viewdirs: bool = True,
                  bf16: bool = False) -> torch.Tensor:
    """REPAIRED PaperNeRFModel.forward (models.py:163-182): the trunk starts from xyz only;
    cat order (xyz, x) at layer 4 (:166-167); fc_feat has no ReLU and alpha = fc_alpha(feat)
    (:171-172); dir branch uses layers_dir[0..2] (:173-180)."""
    dim_xyz = sd["layers_xyz.0.weight"].shape[1]
    xyz, dirs = x[..., :dim_xyz], x[..., dim_xyz:]
    h = xyz
    for i in range(8):
        if i == 4:
            h = torch.cat((xyz, h), dim=-1)
        h = torch.relu(_linear(h, sd, f"layers_xyz.{i}", bf16))
    feat = _linear(h, sd, "fc_feat", bf16)
    alpha = _linear(feat, sd, "fc_alpha", False)
    y = torch.cat((feat, dirs), dim=-1) if use_viewdirs else feat
    for i in range(3):
        y = torch.relu(_linear(y, sd, f"layers_dir.{i}", bf16))
    rgb = _linear(y, sd, "fc_rgb", False)
    return torch.cat((rgb, alpha), dim=-1)


def very_tiny_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """tiny_nerf.py:162-181 / models.py:4-31: relu(l1) -> relu(l2) -> l3."""
    h = torch.relu(_linear(x, sd, "layer1", False))
    h = torch.relu(_linear(h, sd, "layer2", False))
    return _linear(h, sd, "layer3", False)


# --------------------------------------------------------------------------------------
# a-9  orchestration                                          nerf/train_utils.py:72-288
# --------------------------------------------------------------------------------------
class RenderOptions:
    """The subset of the YAML the render path reads."""

    def __init__(self, near, far, num_coarse, num_fine, lindisp=False, white_background=False,
                 noise_std=0.0, perturb=False, use_viewdirs=True, no_ndc=True,
                 Lx=10, Ld=4, include_input_xyz=True, include_input_dir=True,
                 log_sampling_xyz=True, log_sampling_dir=True, chunksize=4096):
        self.__dict__.update(locals())
        del self.__dict__["self"]


def run_network(model_fn, pts: torch.Tensor, viewdirs: Optional[torch.Tensor], o: RenderOptions,
                chunksize: int) -> torch.Tensor:
    """train_utils.py:72-89: encode points (+ broadcast view dirs), chunked MLP, reshape."""
    flat = pts.reshape(-1, 3)
    emb = positional_encoding(flat, o.Lx, o.include_input_xyz, o.log_sampling_xyz)
    if viewdirs is not None:
        d = viewdirs[:, None, :].expand(pts.shape).reshape(-1, 3)
        emb = torch.cat((emb, positional_encoding(d, o.Ld, o.include_input_dir,
                                                  o.log_sampling_dir)), dim=-1)
    out = torch.cat([model_fn(emb[i:i + chunksize]) for i in range(0, emb.shape[0], chunksize)], 0)
    return out.reshape(*pts.shape[:-1], out.shape[-1])


def render_rays(ro: torch.Tensor, rd: torch.Tensor, model_coarse, model_fine, o: RenderOptions,
                m_thres_cand: Sequence[float], t_rand=None, u=None, noise_coarse=None,
                noise_fine=None, height=None, width=None, focal=None, return_aux=False):
    """run_one_iter_of_nerf + predict_and_render_radiance (train_utils.py:92-288) on flat rays
    (n, 3).  Returns (rgb_c, depth_c, acc_c, rgb_f, depth_f, acc_f, *dex_depth_f): slots 1 and 4
    are the EXPECTED DEPTH (train_utils.py:201), not disparity."""
    ro = ro.reshape(-1, 3)
    rd = rd.reshape(-1, 3)
    viewdirs = None
    if o.use_viewdirs:
        rd64 = rd.to(F64)
        nrm = torch.sqrt((rd64 * rd64).sum(-1, keepdim=True)).to(F32)
        viewdirs = rd / nrm
    if not o.no_ndc:
        ro, rd = ndc_rays(height, width, focal, 1.0, ro, rd)
    n = ro.shape[0]
    near = torch.full((n, 1), float(o.near), dtype=F32)
    far = torch.full((n, 1), float(o.far), dtype=F32)
    z = stratified_z(near, far, o.num_coarse, o.lindisp, t_rand)
    pts = ro[:, None, :] + rd[:, None, :] * z[:, :, None]
    rf = run_network(model_coarse, pts, viewdirs, o, o.chunksize)
    c = volume_render_radiance_field(rf, z, rd, o.noise_std, o.white_background, m_thres_cand,
                                     noise=noise_coarse)
    rgb_c, acc_c, w_c, depth_c = c[0], c[2], c[3], c[4]
    if o.num_fine <= 0:
        raise NameError("num_fine == 0 is unsupported by the reference (train_utils.py:201)")
    z_mid = 0.5 * (z[..., 1:] + z[..., :-1])
    z_samples = sample_pdf(z_mid, w_c[..., 1:-1], o.num_fine, det=(not o.perturb), u=u)
    z_samples = z_samples.detach()                       # train_utils.py:170
    z_all = merge_fine(z, z_samples)
    pts = ro[:, None, :] + rd[:, None, :] * z_all[:, :, None]
    rf_f = run_network(model_fine, pts, viewdirs, o, o.chunksize)
    f = volume_render_radiance_field(rf_f, z_all, rd, o.noise_std, o.white_background,
                                     m_thres_cand, noise=noise_fine)
    out = (rgb_c, depth_c, acc_c, f[0], f[4], f[2], *f[5:])
    if return_aux:
        return out, dict(z_coarse=z, weights_coarse=w_c, z_samples=z_samples, z_fine=z_all,
                         rf_coarse=rf, rf_fine=rf_f, weights_fine=f[3], disp_fine=f[1])
    return out


# --------------------------------------------------------------------------------------
# a-10  training iteration                                   train_dexnerf_rgb.py:246-289
# --------------------------------------------------------------------------------------
def img2mse(img_src: torch.Tensor, img_tgt: torch.Tensor) -> torch.Tensor:
    """nerf_helpers.py:9-10 (== torch.nn.functional.mse_loss, mean over all elements)."""
    return ((img_src - img_tgt) ** 2).mean()


def train_loss_and_grads(sd_coarse: Dict[str, torch.Tensor], sd_fine: Dict[str, torch.Tensor],
                         ro: torch.Tensor, rd: torch.Tensor, target: torch.Tensor, o: RenderOptions,
                         m_thres_cand: Sequence[float], skip_coarse: int = 4, skip_fine: int = 4,
                         t_rand=None, u=None, noise_coarse=None, noise_fine=None, bf16: bool = False):
    """One training iteration's loss and parameter gradients (train_dexnerf_rgb.py:246-278):
    train-mode render with the four RNG draws replayed, loss = mse(rgb_coarse, target) +
    mse(rgb_fine, target), autograd backward.  Gradients do not flow through the fine depths
    (train_utils.py:170).  Returns (loss, coarse_loss, fine_loss, grads_coarse, grads_fine)."""
    pc = {k: v.detach().clone().to(F32).requires_grad_(True) for k, v in sd_coarse.items()}
    pf = {k: v.detach().clone().to(F32).requires_grad_(True) for k, v in sd_fine.items()}
    out = render_rays(ro, rd, lambda x: flexible_forward(pc, x, skip_coarse, bf16=bf16),
                      lambda x: flexible_forward(pf, x, skip_fine, bf16=bf16), o, m_thres_cand,
                      t_rand=t_rand, u=u, noise_coarse=noise_coarse, noise_fine=noise_fine)
    coarse_loss = img2mse(out[0][..., :3], target[..., :3])
    fine_loss = img2mse(out[3][..., :3], target[..., :3])
    loss = coarse_loss + fine_loss
    loss.backward()
    gc = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pc.items()}
    gf = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in pf.items()}
    return loss.detach(), coarse_loss.detach(), fine_loss.detach(), gc, gf


def adam_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], lr: float, step: int = 1,
              m=None, v=None, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """torch.optim.Adam's update (train_dexnerf_rgb.py:142-148, defaults) restated; returns new params."""
    out = {}
    for k, p in params.items():
        g = grads[k]
        mk = (1 - beta1) * g if m is None else beta1 * m[k] + (1 - beta1) * g
        vk = (1 - beta2) * g * g if v is None else beta2 * v[k] + (1 - beta2) * g * g
        mhat, vhat = mk / (1 - beta1 ** step), vk / (1 - beta2 ** step)
        out[k] = p - lr * mhat / (vhat.sqrt() + eps)
    return out


def learning_rate(base_lr: float, iteration: int, lr_decay: float, lr_decay_factor: float) -> float:
    """train_dexnerf_rgb.py:283-289."""
    return base_lr * (lr_decay_factor ** (iteration / (lr_decay * 1000)))


# --------------------------------------------------------------------------------------
# next rows: validation depth metrics, camera path
# --------------------------------------------------------------------------------------
def compute_err_metric(depth_gt: torch.Tensor, depth_pred: torch.Tensor, mask: torch.Tensor) -> Dict[str, float]:
    """train_utils.py:9-30: mean |pred - gt| in mm and the fractions of masked pixels whose error
    exceeds 2 / 4 / 8 mm."""
    p, g = depth_pred[mask], depth_gt[mask]
    abs_err = float((p * 1000 - g * 1000).abs().to(F64).mean())
    diff = (g - p).abs()
    n = diff.numel()
    return {"depth_abs_err": abs_err, "depth_err2": float((diff > 2e-3).sum()) / n,
            "depth_err4": float((diff > 4e-3).sum()) / n, "depth_err8": float((diff > 8e-3).sum()) / n}


def select_dex_threshold(depth_planes: Sequence[torch.Tensor], depth_gt: torch.Tensor, mask=None):
    """train_dexnerf_rgb.py:392-404: mask = (gt > 0) & (gt < 1.25) unless given; the first candidate
    with the smallest abs err below 1000 wins.  Returns (index or -1, its error dict or None)."""
    if mask is None:
        mask = (depth_gt > 0) & (depth_gt < 1.25)
    best, best_err, min_abs = -1, None, 1000.0
    for k, plane in enumerate(depth_planes):
        err = compute_err_metric(depth_gt, plane, mask)
        if err["depth_abs_err"] < min_abs:
            min_abs, best, best_err = err["depth_abs_err"], k, err
    return best, best_err


def pose_spherical(theta: float, phi: float, radius: float) -> torch.Tensor:
    """load_blender.py:33-38 (cam->world, OpenGL convention), float32."""
    t = torch.eye(4, dtype=F32); t[2, 3] = radius
    p = phi / 180.0 * math.pi
    rp = torch.eye(4, dtype=F32)
    rp[1, 1] = rp[2, 2] = math.cos(p); rp[1, 2] = -math.sin(p); rp[2, 1] = math.sin(p)
    th = theta / 180.0 * math.pi
    rt = torch.eye(4, dtype=F32)
    rt[0, 0] = rt[2, 2] = math.cos(th); rt[0, 2] = -math.sin(th); rt[2, 0] = math.sin(th)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=F32)
    return flip @ (rt @ (rp @ t))


# --------------------------------------------------------------------------------------
# TinyNeRF (BASELINE config 1)                                      tiny_nerf.py:12-159
# --------------------------------------------------------------------------------------
def tiny_query_points(ro, rd, near: float, far: float, num_samples: int, rand=None):
    """tiny_nerf.py:12-65: a GLOBAL linspace(near, far) plus rand*(far-near)/N jitter."""
    z = torch.linspace(near, far, num_samples, dtype=F32)
    if rand is not None:
        z = z + rand * (far - near) / num_samples
    pts = ro[..., None, :] + rd[..., None, :] * z[..., :, None]
    return pts, z


def tiny_render_volume_density(rf: torch.Tensor, depth_values: torch.Tensor):
    """tiny_nerf.py:68-107: like a-6 without ||rd|| scaling, noise, white bg or Dex depth."""
    sigma = torch.relu(rf[..., 3])
    rgb = torch.sigmoid(rf[..., :3])
    z = depth_values
    dists = torch.cat((z[..., 1:] - z[..., :-1], torch.full_like(z[..., :1], 1e10)), dim=-1)
    alpha = 1.0 - torch.exp(-sigma * dists)
    w = alpha * cumprod_exclusive(1.0 - alpha + 1e-10)
    w64 = w.to(F64)
    rgb_map = (w64[..., None] * rgb.to(F64)).sum(-2).to(F32)
    depth_map = (w64 * z.to(F64)).sum(-1).to(F32)
    acc_map = w64.sum(-1).to(F32)
    return rgb_map, depth_map, acc_map


def run_one_iter_of_tinynerf(height, width, tform, intrinsic, near, far, num_samples, L,
                             model_fn, chunksize=16384):
    """tiny_nerf.py:111-159 with the 5-argument get_ray_bundle."""
    ro, rd = get_ray_bundle(height, width, None, tform, intrinsic)
    pts, z = tiny_query_points(ro, rd, near, far, num_samples)
    flat = pts.reshape(-1, 3)
    enc = positional_encoding(flat, L)
    rf = torch.cat([model_fn(enc[i:i + chunksize]) for i in range(0, enc.shape[0], chunksize)], 0)
    rf = rf.reshape(*pts.shape[:-1], 4)
    return tiny_render_volume_density(rf, z)[0]


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic workloads of SURVEY.md section 8d)
# --------------------------------------------------------------------------------------
def flexible_shapes(num_layers=8, hidden=256, skip=4, Lx=10, Ld=4, include_xyz=True,
                    include_dir=True) -> List[Tuple[str, int, int]]:
    """(name, out_features, in_features) in torch construction order (models.py:197-229)."""
    dx = (3 if include_xyz else 0) + 6 * Lx
    dd = (3 if include_dir else 0) + 6 * Ld
    shapes = [("layer1", hidden, dx)]
    for i in range(num_layers - 1):
        k = dx + hidden if (i % skip == 0 and i > 0 and i != num_layers - 1) else hidden
        shapes.append((f"layers_xyz.{i}", hidden, k))
    shapes.append(("layers_dir.0", hidden // 2, dd + hidden))
    shapes.append(("fc_alpha", 1, hidden))
    shapes.append(("fc_rgb", 3, hidden // 2))
    shapes.append(("fc_feat", hidden, hidden))
    return shapes


def init_flexible_state_dict(seed_generator: torch.Generator, **kw) -> Dict[str, torch.Tensor]:
    """torch.nn.Linear's default init (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in))
    for weight and bias) drawn from an explicit generator, in construction order."""
    sd = {}
    for name, out_f, in_f in flexible_shapes(**kw):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=seed_generator, dtype=F32) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(out_f, generator=seed_generator, dtype=F32) * 2 - 1) * bound
    return sd


def pose_spherical_world2cam(theta_deg: float, phi_deg: float, radius: float) -> torch.Tensor:
    """Camera on a sphere looking at the origin, returned as the world->cam 4x4 the fork's
    get_ray_bundle expects (load_blender.py:33-38 builds cam->world; we invert it)."""
    def trans_t(t):
        m = torch.eye(4, dtype=F64); m[2, 3] = t; return m

    def rot_phi(p):
        m = torch.eye(4, dtype=F64)
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = math.cos(p), -math.sin(p), math.sin(p), math.cos(p)
        return m

    def rot_theta(t):
        m = torch.eye(4, dtype=F64)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = math.cos(t), -math.sin(t), math.sin(t), math.cos(t)
        return m

    c2w = trans_t(radius)
    c2w = rot_phi(phi_deg / 180.0 * math.pi) @ c2w
    c2w = rot_theta(theta_deg / 180.0 * math.pi) @ c2w
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=F64)
    c2w = flip @ c2w
    # OpenGL camera (looks along -z, y up) -> the fork's OpenCV camera (rays along +z, y down)
    c2w = c2w @ torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], dtype=F64))
    return torch.linalg.inv(c2w).to(F32)
