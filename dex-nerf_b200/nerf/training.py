"""Training step of the render hot path (reference: train_dexnerf_rgb.py:246-289 -
run_one_iter_of_nerf(mode="train"), mse(rgb_coarse) + mse(rgb_fine), loss.backward(), Adam,
exponential learning-rate decay).

`FieldRender` is one autograd node per network pass: forward = fused tensor-core query that also
records the training tape + compositing; backward = compositing backward -> activation-gradient
chain -> weight-gradient GEMM, all hand-written sm_100a kernels behind the C ABI.  Gradients flow
to the model parameters only - depths are detached exactly where the reference detaches them
(train_utils.py:170) and ray origins / directions never require grad in the reference scripts."""
import ctypes as C

import torch

from . import _lib as L
from . import tensorcore
from .volume_rendering_utils import render_maps


event_log = None     # bench.py sets this to a list to collect (name, start_event, end_event, n, S)


class _timed:
    """CUDA-event bracket around one C-ABI call when bench.py asked for a kernel breakdown."""

    def __init__(self, name, n, S):
        self.name, self.n, self.S = name, n, S

    def __enter__(self):
        if event_log is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if event_log is not None:
            self.e1.record()
            event_log.append((self.name, self.e0, self.e1, self.n, self.S))


def _layer_params(model):
    return [p for lin, *_ in model._layers() for p in (lin.weight, lin.bias)]


def tape_layout(spec, n_samples):
    """Offsets (bytes) of the tape images; see csrc/tc_plan.cuh TapeLayout."""
    out = (C.c_int64 * 54)()
    L.check(L.lib().dexnerf_tc_tape_layout(spec, int(n_samples), out), "tc_tape_layout")
    L.launch_count -= 1          # host-only call, not a kernel launch
    nl = int(out[0])
    return dict(nl=nl, n_tiles=int(out[1]), total=int(out[2]), ghead=int(out[3]), xyz=int(out[4]), dir=int(out[5]),
                act=[int(out[6 + l]) for l in range(nl)], mask=[int(out[22 + l]) for l in range(nl)],
                grad=[int(out[38 + l]) for l in range(nl)])


def decode_image(tape, offset, n_tiles, features):
    """A tape image -> (n_tiles * 128, features) float32 (test / debugging aid)."""
    nbytes = n_tiles * features * 256
    img = tape[offset:offset + nbytes].view(torch.bfloat16).view(n_tiles, 2, features // 8, 64, 8)
    return img.permute(0, 1, 3, 2, 4).reshape(n_tiles * 128, features).float()


def encode_image(tape, offset, n_tiles, x):
    """Inverse of decode_image: write (n_tiles * 128, features) values into the tape as bf16."""
    features = x.shape[1]
    img = x.to(torch.bfloat16).view(n_tiles, 2, 64, features // 8, 8).permute(0, 1, 3, 2, 4).contiguous()
    tape[offset:offset + n_tiles * features * 256] = img.view(torch.uint8).reshape(-1)


def packed_weights_t(model, prog, spec):
    """Transposed bf16 weight images for the activation-gradient chain, cached per parameter version."""
    params = model.packed_params()
    key = (params.data_ptr(), model.__dict__["_packed_cache"][0])
    cache = model.__dict__.get("_tc_cache_t")
    if cache is None or cache[0] != key:
        nbytes = L.lib().dexnerf_tc_packed_bwd_bytes(spec)
        if nbytes <= 0:
            raise L.DexNerfError("tc_packed_bwd_bytes: " + L.lib().dexnerf_last_error().decode())
        blob = torch.empty(nbytes, dtype=torch.uint8, device=params.device)
        L.check(L.lib().dexnerf_tc_pack_bwd(spec, prog, L.ptr(params), L.ptr(blob), L.stream_ptr()), "tc_pack_bwd")
        cache = (key, blob)
        model.__dict__["_tc_cache_t"] = cache
    return cache[1]


def query_train(model, prog, spec, ro, rd, viewdirs, z):
    """Forward query that records the tape.  Returns (rf (n,S,4), tape uint8)."""
    n, S = z.shape
    blob = tensorcore.packed_weights(model, prog, spec)
    nbytes = L.lib().dexnerf_tc_tape_bytes(spec, n * S)
    if nbytes < 0:
        raise L.DexNerfError("tc_tape_bytes: " + L.lib().dexnerf_last_error().decode())
    tape = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
    rf = torch.empty((n, S, 4), dtype=torch.float32, device=z.device)
    with _timed("mlp_tc_train_fwd", n, S):
        L.check(L.lib().dexnerf_tc_query_train(spec, L.ptr(blob), L.ptr(ro), L.ptr(rd), L.ptr(viewdirs), L.ptr(z), n,
                                               S, L.ptr(rf), L.ptr(tape), L.stream_ptr()), "tc_query_train")
    return rf, tape


def mlp_backward(model, prog, spec, tape, d_rf, n, S, what=3, variant=0):
    """dL/d(rf) -> flat fp32 gradient buffer in the layout of model.packed_params()."""
    blob = tensorcore.packed_weights(model, prog, spec)
    blob_t = packed_weights_t(model, prog, spec)
    flat = torch.zeros_like(model.packed_params())
    for bit, name in ((1, "mlp_tc_bwd_dx"), (2, "mlp_tc_bwd_dw")):     # one kernel per call
        if what & bit:
            with _timed(name, n, S):
                L.check(L.lib().dexnerf_tc_backward(spec, prog, L.ptr(blob), L.ptr(blob_t), L.ptr(tape), L.ptr(d_rf),
                                                    n, S, L.ptr(flat), bit, int(variant), L.stream_ptr()),
                        "tc_backward")
    return flat


def unflatten_grads(model, prog, flat):
    """Program-layout gradient buffer -> [dW (out,in), db] per nn.Linear in _layers() order."""
    out = []
    for i, (lin, *_rest) in enumerate(model._layers()):
        op = prog.ops[i]
        fin, fout = lin.in_features, lin.out_features
        out.append(flat[op.w_off:op.w_off + fin * fout].view(fin, fout).t())
        out.append(flat[op.b_off:op.b_off + fout])
    return out


def volume_render_backward(rf, z, rd, noise, white_background, g_rgb, g_depth, g_acc):
    n, S = z.shape
    d_rf = torch.empty_like(rf)
    if n:
        g = [None if t is None else t.contiguous().to(torch.float32) for t in (g_rgb, g_depth, g_acc)]
        with _timed("composite_bwd", n, S):
            L.check(L.lib().dexnerf_volume_render_backward(L.ptr(rf), L.ptr(z), L.ptr(rd), L.ptr(noise), n, S,
                                                           int(bool(white_background)), L.ptr(g[0]), L.ptr(g[1]),
                                                           L.ptr(g[2]), L.ptr(d_rf), L.stream_ptr()),
                    "volume_render_backward")
    return d_rf


class FieldRender(torch.autograd.Function):
    """(model, rays, depths) -> (rgb_map, depth_map, acc_map, weights, dex depths) with gradients
    to the model parameters.  weights and the Dex depths are returned detached (the reference only
    consumes them detached / for logging)."""

    @staticmethod
    def forward(ctx, model, embed_fn, embeddirs_fn, ro, rd, viewdirs, z, noise, white_background, thr, T,
                *params):
        prog = model.program(embed_fn, embeddirs_fn)
        spec = tensorcore.spec_for(model, prog)
        if spec is None:
            raise L.DexNerfError("training runs on the tensor-core path: FlexibleNeRFModel with view directions, "
                                 "hidden size 128 or 256")
        n, S = z.shape
        rf, tape = query_train(model, prog, spec, ro, rd, viewdirs, z)
        o = render_maps(rf, z, rd, noise, white_background, thr, T)
        ctx.model, ctx.prog, ctx.spec = model, prog, spec
        ctx.white = bool(white_background)
        ctx.noise = noise
        ctx.save_for_backward(rf, z, rd, tape)
        dex = o["dex"] if T else torch.empty((0, n), device=z.device)
        ctx.mark_non_differentiable(o["weights"], dex)
        return o["rgb"], o["depth"], o["acc"], o["weights"], dex

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_acc, _g_w, _g_dex):
        rf, z, rd, tape = ctx.saved_tensors
        n, S = z.shape
        d_rf = volume_render_backward(rf, z, rd, ctx.noise, ctx.white, g_rgb, g_depth, g_acc)
        flat = mlp_backward(ctx.model, ctx.prog, ctx.spec, tape, d_rf, n, S)
        grads = unflatten_grads(ctx.model, ctx.prog, flat)
        return (None,) * 11 + tuple(grads)


def render_field(model, embed_fn, embeddirs_fn, ro, rd, viewdirs, z, noise, white_background, thr, T):
    params = _layer_params(model)
    return FieldRender.apply(model, embed_fn, embeddirs_fn, ro, rd, viewdirs, z, noise, white_background, thr, T,
                             *params)


def wants_grad(*models):
    return torch.is_grad_enabled() and any(p.requires_grad for m in models if m is not None for p in m.parameters())


def learning_rate(base_lr, iteration, lr_decay, lr_decay_factor):
    """train_dexnerf_rgb.py:283-289: lr * factor ** (i / (lr_decay * 1000))."""
    return base_lr * (lr_decay_factor ** (iteration / (lr_decay * 1000)))


def train_step(model_coarse, model_fine, optimizer, ro, rd, target, cfg, encode_position_fn, encode_direction_fn,
               m_thres_cand=(), rng=None, height=0, width=0, focal=1.0, world_size=1):
    """One reference training iteration on pre-selected rays (train_dexnerf_rgb.py:246-281).
    Returns (loss, psnr-ready mse terms).  With world_size > 1 the caller all-reduces the gradients
    (nerf.allreduce_gradients) between backward and optimizer.step()."""
    from .train_utils import run_one_iter_of_nerf
    out = run_one_iter_of_nerf(height, width, focal, model_coarse, model_fine, ro, rd, cfg, mode="train",
                               encode_position_fn=encode_position_fn, encode_direction_fn=encode_direction_fn,
                               m_thres_cand=list(m_thres_cand), rng=rng)
    coarse_loss = torch.nn.functional.mse_loss(out[0][..., :3], target[..., :3])
    fine_loss = torch.nn.functional.mse_loss(out[3][..., :3], target[..., :3])
    loss = coarse_loss + fine_loss
    loss.backward()
    if world_size > 1:
        from .sharding import allreduce_gradients
        allreduce_gradients([model_coarse, model_fine])
    optimizer.step()
    optimizer.zero_grad()
    return loss.detach(), coarse_loss.detach(), fine_loss.detach()
