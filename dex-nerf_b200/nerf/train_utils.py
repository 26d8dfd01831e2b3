"""Drop-in for the render driver of the reference's nerf/train_utils.py:72-288
(run_network, predict_and_render_radiance, run_one_iter_of_nerf).

The reference evaluates a ray chunk as ~200 eager ATen kernels and materialises the positional
encoding of every sample (10 GB for an 800x800 frame).  Here a chunk is six launches
(coarse depths, coarse MLP query, coarse compositing, resample+merge, fine MLP query, fine
compositing); encodings never leave the SM.  Return tuples, shapes and quirks follow the reference.
"""
import os

import torch

from . import _lib as L
from .nerf_helpers import _Embedder, get_minibatches, ndc_rays
from .nerf_helpers import sample_pdf_2 as sample_pdf  # the reference re-binds the name (train_utils.py:6)
from .volume_rendering_utils import _thresholds_tensor, render_maps, volume_render_radiance_field

# "bf16": tcgen05 tensor-core MLP (bf16 operands, fp32 accumulate) where the model is supported;
# "fp32": CUDA-core MLP with the reference's fp32 arithmetic.  DEXNERF_PRECISION overrides.
_precision = os.environ.get("DEXNERF_PRECISION", "bf16")


def set_precision(p):
    global _precision
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _precision = p


def get_precision():
    return _precision


def _fusable(model, embed_fn, embeddirs_fn):
    return (hasattr(model, "program") and isinstance(embed_fn, _Embedder)
            and (embeddirs_fn is None or isinstance(embeddirs_fn, _Embedder))
            and embed_fn.out_dim == model.dim_xyz
            and (model.dim_dir == 0 or (embeddirs_fn is not None and embeddirs_fn.out_dim == model.dim_dir)))


def _trainable(model, embed_fn, embeddirs_fn):
    from . import tensorcore
    return (_fusable(model, embed_fn, embeddirs_fn)
            and tensorcore.trainable(model, model.program(embed_fn, embeddirs_fn if model.dim_dir else None)))


def _no_grad_reason(model, embed_fn, embeddirs_fn):
    """Why `model` cannot be trained through this package (None when it can)."""
    from .models import FlexibleNeRFModel
    if _precision != "bf16":
        return "precision is %r (set_precision('bf16') / DEXNERF_PRECISION selects the tensor-core path)" % _precision
    if not isinstance(model, FlexibleNeRFModel):
        return "%s has no backward kernels (only FlexibleNeRFModel does)" % type(model).__name__
    if not model.use_viewdirs:
        return "FlexibleNeRFModel(use_viewdirs=False) has no backward kernels"
    if model.hidden_size not in (128, 256):
        return "hidden_size %d (the tensor-core kernels exist for 128 and 256)" % model.hidden_size
    if not isinstance(embed_fn, _Embedder) or not isinstance(embeddirs_fn, _Embedder):
        return "the encoders are not get_embedding_function objects (a custom encode fn cannot be fused)"
    if model.dim_xyz > 64 or model.dim_dir > 32:
        return "encoding widths %d / %d exceed the kernel's 64 / 32" % (model.dim_xyz, model.dim_dir)
    if not _trainable(model, embed_fn, embeddirs_fn):
        return "the model / encoder combination is not supported by the tensor-core training kernels"
    return None


kernel_event_log = None   # bench.py sets this to a list to collect (name, start_evt, end_evt, n, S)


def query_field(model, ro, rd, viewdirs, z, embed_fn, embeddirs_fn):
    """pts = ro + rd*z -> encode -> model, fused in one kernel.  ro, rd, viewdirs (n,3), z (n,S)
    -> raw radiance field (n,S,4)."""
    n, S = z.shape
    rf = torch.empty((n, S, 4), dtype=torch.float32, device=z.device)
    if n == 0:
        return rf
    prog = model.program(embed_fn, embeddirs_fn if model.dim_dir else None)
    vd = viewdirs if model.dim_dir else None
    log = kernel_event_log
    if log is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    used = "mlp_simt"
    if _precision == "bf16" and _tc_query(model, prog, ro, rd, vd, z, rf):
        used = "mlp_tc"
    else:
        L.check(L.lib().dexnerf_mlp_query(prog, L.ptr(model.packed_params()), L.ptr(ro), L.ptr(rd), L.ptr(vd),
                                          L.ptr(z), n, S, L.ptr(rf), L.stream_ptr()), "mlp_query")
    if log is not None:
        e1.record()
        log.append((used, e0, e1, n, S))
    return rf


def _tc_query(model, prog, ro, rd, vd, z, rf):
    from . import tensorcore
    if not tensorcore.supported(model, prog):
        return False
    tensorcore.query(model, prog, ro, rd, vd, z, rf)
    return True


def run_network(network_fn, pts, ray_batch, chunksize, embed_fn, embeddirs_fn):
    """train_utils.py:72-89, for callers that hand in explicit points: encode, append the encoded
    view directions (last three ray columns, :77), chunked model calls, reshape."""
    pts_flat = pts.reshape((-1, pts.shape[-1]))
    embedded = embed_fn(pts_flat)
    if embeddirs_fn is not None:
        viewdirs = ray_batch[..., None, -3:]
        input_dirs_flat = viewdirs.expand(pts.shape).reshape((-1, 3))
        embedded = torch.cat((embedded, embeddirs_fn(input_dirs_flat.contiguous())), dim=-1)
    preds = [network_fn(batch) for batch in get_minibatches(embedded, chunksize=chunksize)]
    radiance_field = torch.cat(preds, dim=0)
    return radiance_field.reshape(list(pts.shape[:-1]) + [radiance_field.shape[-1]])


def _coarse_depths(n, near, far, num_coarse, lindisp, t_rand, device):
    z = torch.empty((n, num_coarse), dtype=torch.float32, device=device)
    if n:
        L.check(L.lib().dexnerf_stratified_z(n, num_coarse, 0.0, 0.0, L.ptr(near), L.ptr(far), int(bool(lindisp)),
                                             L.ptr(t_rand), L.ptr(z), L.stream_ptr()), "stratified_z")
    return z


def _field(model, ro, rd, viewdirs, z, ray_batch, chunksize, embed_fn, embeddirs_fn):
    if _fusable(model, embed_fn, embeddirs_fn):
        return query_field(model, ro, rd, viewdirs, z, embed_fn, embeddirs_fn)
    # arbitrary callables: same dataflow as the reference, each step one of our kernels
    pts = ro[..., None, :] + rd[..., None, :] * z[..., :, None]
    return run_network(model, pts, ray_batch, chunksize, embed_fn, embeddirs_fn).contiguous()


def predict_and_render_radiance(ray_batch, model_coarse, model_fine, options, mode="train",
                                encode_position_fn=None, encode_direction_fn=None, m_thres_cand=None,
                                rng=None):
    """train_utils.py:92-202.  ray_batch (n, 8 | 11) = [ro, rd, near, far, (viewdir)].
    Returns (rgb_coarse, depth_coarse, acc_coarse, rgb_fine, depth_fine, acc_fine, *depth_fine_dex).
    `rng` (extension): dict with any of t_rand (n,Nc), u (n,Nf), noise_coarse (n,Nc),
    noise_fine (n,Nc+Nf) to replay the reference's four random draws."""
    rng = rng or {}
    opt = getattr(options.nerf, mode)
    rays = L.dev_f32(ray_batch, "ray_batch")
    n = rays.shape[0]
    dev = rays.device
    ro, rd = rays[:, 0:3].contiguous(), rays[:, 3:6].contiguous()
    near, far = rays[:, 6].contiguous(), rays[:, 7].contiguous()
    viewdirs = rays[:, -3:].contiguous() if rays.shape[1] > 8 else None
    thr, T = _thresholds_tensor(m_thres_cand, dev)
    Nc, Nf = int(opt.num_coarse), int(opt.num_fine)
    std = float(opt.radiance_field_noise_std)

    t_rand = None
    if opt.perturb:
        t_rand = rng.get("t_rand")
        if t_rand is None:
            t_rand = torch.rand((n, Nc), dtype=torch.float32, device=dev)
        t_rand = L.dev_f32(t_rand, "t_rand")
    z = _coarse_depths(n, near, far, Nc, opt.lindisp, t_rand, dev)

    from . import training
    # Gradients are recorded in "train" mode only (the reference's scripts render validation frames
    # under torch.no_grad(), train_dexnerf_rgb.py:317; a full frame's tape would not fit anyway).
    with_grad = n > 0 and mode == "train" and training.wants_grad(model_coarse, model_fine)
    if with_grad:
        # No silent outputs without a grad_fn (loss.backward() would fail later with an opaque autograd error):
        # gradients exist on the tensor-core path only - say so here, naming the condition.
        for m in (model_coarse, model_fine):
            why = _no_grad_reason(m, encode_position_fn, encode_direction_fn)
            if why is not None:
                raise L.DexNerfError("run_one_iter_of_nerf(mode='train') with gradients enabled is not supported for "
                                     "this configuration: " + why + ".  Gradients are implemented for "
                                     "FlexibleNeRFModel with view directions, hidden 128/256, "
                                     "get_embedding_function encoders, precision 'bf16'; wrap the call in "
                                     "torch.no_grad() to render without gradients")
    noise = rng.get("noise_coarse")
    if noise is None and std > 0.0:
        noise = torch.randn((n, Nc), dtype=torch.float32, device=dev) * std
    noise = L.dev_f32(noise, "noise_coarse", allow_none=True)
    if with_grad:
        # training: one autograd node per network pass (fused query + tape, compositing; backward =
        # compositing backward + tensor-core activation / weight gradient kernels)
        rgb_coarse, depth_coarse, acc_coarse, w_coarse, _ = training.render_field(
            model_coarse, encode_position_fn, encode_direction_fn, ro, rd, viewdirs, z, noise,
            opt.white_background, thr, T)
        c = dict(weights=w_coarse)
    else:
        rf = _field(model_coarse, ro, rd, viewdirs, z, rays, opt.chunksize, encode_position_fn, encode_direction_fn)
        c = render_maps(rf, z, rd, noise, opt.white_background, thr, T)
        rgb_coarse, acc_coarse, depth_coarse = c["rgb"], c["acc"], c["depth"]

    if Nf <= 0:
        # the reference builds its return from depth_fine_dex, which only exists after the fine pass
        raise NameError("name 'depth_fine_dex' is not defined")
    u = rng.get("u")
    if u is None and opt.perturb != 0.0:       # det = (perturb == 0.0), train_utils.py:169
        u = torch.rand((n, Nf), dtype=torch.float32, device=dev)
    u = L.dev_f32(u, "u", allow_none=True)
    z_fine = torch.empty((n, Nc + Nf), dtype=torch.float32, device=dev)
    if n:
        with L.timed("resample_merge", n, Nc + Nf, n * (8 * Nc + 4 * (Nc + Nf) + (4 * Nf if u is not None else 0))):
            L.check(L.lib().dexnerf_resample_merge(L.ptr(z), L.ptr(c["weights"]), n, Nc, Nf, L.ptr(u), L.ptr(z_fine),
                                                   L.stream_ptr()), "resample_merge")
    noise = rng.get("noise_fine")
    if noise is None and std > 0.0:
        noise = torch.randn((n, Nc + Nf), dtype=torch.float32, device=dev) * std
    noise = L.dev_f32(noise, "noise_fine", allow_none=True)
    if with_grad:
        rgb_f, depth_f, acc_f, _, dex_f = training.render_field(
            model_fine, encode_position_fn, encode_direction_fn, ro, rd, viewdirs, z_fine, noise,
            opt.white_background, thr, T)
        return tuple([rgb_coarse, depth_coarse, acc_coarse, rgb_f, depth_f, acc_f] + [dex_f[t] for t in range(T)])
    rf_f = _field(model_fine, ro, rd, viewdirs, z_fine, rays, opt.chunksize, encode_position_fn,
                  encode_direction_fn)
    f = render_maps(rf_f, z_fine, rd, noise, opt.white_background, thr, T, want_weights=False)
    dex = [f["dex"][t] for t in range(T)]
    return tuple([rgb_coarse, depth_coarse, acc_coarse, f["rgb"], f["depth"], f["acc"]] + dex)


def _fused_ok(model_coarse, model_fine, mode, embed_fn, embeddirs_fn):
    """The fused driver (csrc/render.cu) runs every call whose two models lower to layer programs with
    get_embedding_function encoders, except a training call that has to record an autograd graph."""
    from . import training
    if model_coarse is None or model_fine is None:
        return False
    if not (_fusable(model_coarse, embed_fn, embeddirs_fn) and _fusable(model_fine, embed_fn, embeddirs_fn)):
        return False
    return not (mode == "train" and training.wants_grad(model_coarse, model_fine))


def _run_fused(height, width, focal_length, model_coarse, model_fine, ro_in, rd_in, options, mode, embed_fn,
               embeddirs_fn, m_thres_cand, rng, camera=None, out=None):
    """run_one_iter_of_nerf as one C-ABI call per ray chunk (6 launches; no torch kernel)."""
    from . import render
    opt = getattr(options.nerf, mode)
    dev = rd_in.device if rd_in is not None else camera[0].device
    thr, T = _thresholds_tensor(m_thres_cand, dev)
    if int(opt.num_fine) <= 0:
        # the reference builds its return from depth_fine_dex, which only exists after the fine pass
        raise NameError("name 'depth_fine_dex' is not defined")
    if camera is None:
        shape = tuple(rd_in.shape)
        ro, rd = ro_in.reshape((-1, 3)), rd_in.reshape((-1, 3))
    else:
        shape = (int(camera[3]), int(width), 3)
        ro = rd = None
    if (camera is None and rd.shape[0] == 0) or (camera is not None and shape[0] * shape[1] == 0):
        return ()      # no rays -> no ray chunks -> the reference's zip(*[]) is empty (train_utils.py:252-282)
    outs = render.render_rays(height, width, focal_length, model_coarse, model_fine, ro, rd, options, mode, embed_fn,
                              embeddirs_fn, thr, T, rng, _precision, camera=camera, out=out)
    if out is not None and T:      # the Dex planes live in the shared frame, out[4] floats apart
        n = outs[3].shape[0]
        images = outs[:6] + [torch.as_strided(outs[6], (n,), (1,), outs[6].storage_offset() + t * int(out[4]))
                             for t in range(T)]
    else:
        images = outs[:6] + [outs[6][t] for t in range(T)]
    if mode == "validation":
        shapes = [shape, shape[:-1], shape[:-1]] * 2 + [shape[:-1]] * T
        images = [image.view(s) for image, s in zip(images, shapes)]
    return tuple(images)


def render_camera(height, width, tform_cam2world, intrinsic, model_coarse, model_fine, options, mode="validation",
                  encode_position_fn=None, encode_direction_fn=None, m_thres_cand=None, row_start=0, row_count=None,
                  rng=None, frame=None):
    """get_ray_bundle + run_one_iter_of_nerf for one camera in ONE fused call per chunk (extension): the rays of
    image rows [row_start, row_start + row_count) are generated inside the setup launch from the world->cam
    extrinsic and the intrinsic (same quirks as get_ray_bundle, nerf_helpers.py:67-112), so a frame is
    6 launches.  Returns run_one_iter_of_nerf's tuple, shaped (rows, width[, 3]) in validation mode.
    `frame` (a nerf.SharedFrame): the fine pass's planes of these rows are written straight into the frame held by
    the root rank - over NVLink from the compositing kernel itself - instead of into local tensors."""
    from .nerf_helpers import _small_to_device
    T = _small_to_device(tform_cam2world, "tform_cam2world", (4, 4))
    K = _small_to_device(intrinsic, "intrinsic", (3, 3))
    rows = int(height) - int(row_start) if row_count is None else int(row_count)
    if not _fused_ok(model_coarse, model_fine, mode, encode_position_fn, encode_direction_fn):
        if frame is not None:
            raise L.DexNerfError("render_camera(frame=...) needs the fused render path (tensor-core or fp32 kernel models "
                                 "with nerf.get_embedding_function encoders)")
        from .nerf_helpers import get_ray_bundle
        ro, rd = get_ray_bundle(height, width, None, T, K, row_start=row_start, row_count=rows)
        return run_one_iter_of_nerf(height, width, K[0, 0], model_coarse, model_fine, ro, rd, options, mode=mode,
                                    encode_position_fn=encode_position_fn, encode_direction_fn=encode_direction_fn,
                                    m_thres_cand=m_thres_cand, rng=rng)
    focal = float(K[0, 0]) if options.dataset.no_ndc is False else 0.0     # only ndc_rays reads it (a host sync)
    out = None
    if frame is not None:
        if (int(height), int(width)) != (frame.H, frame.W) or len(list(m_thres_cand)) != frame.T:
            raise ValueError("render_camera: the SharedFrame was made for another frame size / threshold count")
        out = frame.outputs(row_start, rows)
    return _run_fused(height, width, focal, model_coarse, model_fine, None, None, options, mode, encode_position_fn,
                      encode_direction_fn, m_thres_cand, rng, camera=(T, K, int(row_start), rows), out=out)


def run_one_iter_of_nerf(height, width, focal_length, model_coarse, model_fine, ray_origins, ray_directions,
                         options, mode="train", encode_position_fn=None, encode_direction_fn=None,
                         m_thres_cand=None, rng=None):
    """train_utils.py:205-288.  Slots 1 and 4 of the result are the EXPECTED depth (:201), the
    Dex-NeRF threshold depths of the fine pass follow from slot 6.  In "validation" mode outputs
    are reshaped back to the shape of `ray_directions`.  The ray chunk size of the YAML
    (`chunksize`) still bounds the rays per launch; the per-sample chunking of the reference's
    run_network is unnecessary because no per-sample tensor but (r,g,b,sigma) exists."""
    rd_in = L.dev_f32(ray_directions, "ray_directions")
    ro_in = L.dev_f32(ray_origins, "ray_origins")
    if _fused_ok(model_coarse, model_fine, mode, encode_position_fn, encode_direction_fn):
        return _run_fused(height, width, focal_length, model_coarse, model_fine, ro_in, rd_in, options, mode,
                          encode_position_fn, encode_direction_fn, m_thres_cand, rng)
    viewdirs = None
    if options.nerf.use_viewdirs:
        viewdirs = rd_in / rd_in.norm(p=2, dim=-1).unsqueeze(-1)
        viewdirs = viewdirs.reshape((-1, 3))
    restore_shapes = [rd_in.shape, rd_in.shape[:-1], rd_in.shape[:-1]]
    if model_fine:
        restore_shapes += restore_shapes
        for _ in m_thres_cand:
            restore_shapes += [rd_in.shape[:-1]]
    if options.dataset.no_ndc is False:
        ro, rd = ndc_rays(height, width, focal_length, 1.0, ro_in, rd_in)
        ro, rd = ro.reshape((-1, 3)), rd.reshape((-1, 3))
    else:
        ro, rd = ro_in.reshape((-1, 3)), rd_in.reshape((-1, 3))
    near = options.dataset.near * torch.ones_like(rd[..., :1])
    far = options.dataset.far * torch.ones_like(rd[..., :1])
    rays = torch.cat((ro, rd, near, far), dim=-1)
    if options.nerf.use_viewdirs:
        rays = torch.cat((rays, viewdirs), dim=-1)

    chunk = getattr(options.nerf, mode).chunksize
    pred = []
    for start in range(0, rays.shape[0], chunk):
        sub = None
        if rng:
            sub = {k: v[start:start + chunk] for k, v in rng.items() if v is not None}
        pred.append(predict_and_render_radiance(rays[start:start + chunk], model_coarse, model_fine, options,
                                                mode=mode, encode_position_fn=encode_position_fn,
                                                encode_direction_fn=encode_direction_fn,
                                                m_thres_cand=m_thres_cand, rng=sub))
    if len(pred) == 1:
        synthesized_images = list(pred[0])
    else:
        synthesized_images = [torch.cat(image, dim=0) if image[0] is not None else None
                              for image in zip(*pred)]
    if mode == "validation":
        synthesized_images = [image.view(shape) if image is not None else None
                              for (image, shape) in zip(synthesized_images, restore_shapes)]
        if model_fine:
            return tuple(synthesized_images)
        return tuple(synthesized_images + [None, None, None])
    return tuple(synthesized_images)
