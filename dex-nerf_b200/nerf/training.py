"""Training step of the render hot path (reference: train_dexnerf_rgb.py:246-289 -
run_one_iter_of_nerf(mode="train"), mse(rgb_coarse) + mse(rgb_fine), loss.backward(), Adam,
exponential learning-rate decay).

`FieldRender` is one autograd node per network pass: forward = fused tensor-core query that also
records the training tape + compositing; backward = compositing backward -> activation-gradient
chain -> weight-gradient GEMM, all hand-written sm_100a kernels behind the C ABI.  Gradients flow
to the model parameters only - depths are detached exactly where the reference detaches them
(train_utils.py:170) and ray origins / directions never require grad in the reference scripts."""
import ctypes as C

import os

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
    if what in (4, 8):     # the fused launches (chain + weight-gradient GEMM concurrently, G through L2)
        with _timed("mlp_tc_bwd_fused", n, S):
            L.check(L.lib().dexnerf_tc_backward(spec, prog, L.ptr(blob), L.ptr(blob_t), L.ptr(tape), L.ptr(d_rf),
                                                n, S, L.ptr(flat), what, int(variant), L.stream_ptr()), "tc_backward")
        return flat
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
        spec = tensorcore.spec_for(model, prog) if tensorcore.trainable(model, prog) else None
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


# ---------------------------------------------------------------------------------------------------
# Trainer: the training iteration without per-tensor glue
# ---------------------------------------------------------------------------------------------------
class _DeviceMemory:
    """A raw device allocation as a __cuda_array_interface__ object (float32 vector)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def _wrap_device_memory(ptr, n, dev):
    return torch.as_tensor(_DeviceMemory(ptr, n), device=dev)


class Trainer:
    """The reference's training iteration (train_dexnerf_rgb.py:246-289) on FLAT buffers.

    Parameters of both networks live in one fp32 buffer in the kernels' program layout
    (`[coarse | fine]`, each `Wt[in][out] | bias` per layer); gradients, Adam's two moment buffers
    and the NCCL all-reduce use the same layout, so one iteration is ~35 launches: depths, two
    forward queries with tape, two compositings, resampling, loss, two compositing backwards, two
    activation-gradient chains, two weight-gradient GEMMs (reducing straight into the flat gradient
    buffer), one all-reduce, ONE fused Adam step, and the re-packing of the bf16 weight images.
    The arithmetic is identical to the autograd path (`run_one_iter_of_nerf(mode="train")` +
    `torch.optim.Adam`); `sync_to_modules()` writes the parameters back into the nn.Modules (for
    checkpoints with the reference's state-dict keys, or for rendering through the drop-in API)."""

    def __init__(self, model_coarse, model_fine, options, encode_position_fn, encode_direction_fn, lr=5e-3,
                 betas=(0.9, 0.999), eps=1e-8, lr_decay=250, lr_decay_factor=0.1, world_size=1, group=None):
        self.models = (model_coarse, model_fine)
        self.cfg = options
        self.ex, self.ed = encode_position_fn, encode_direction_fn
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.lr_decay, self.lr_decay_factor = lr_decay, lr_decay_factor
        self.world, self.group = int(world_size), group
        # `iteration` is the loop index of the NEXT step (the reference's `i`), `adam_steps` the number of Adam
        # updates made so far and `_lr_now` the rate the optimizer currently holds.  They only differ after a
        # resume: the reference re-runs loop index `iter` of its checkpoint (train_dexnerf_rgb.py:172-178) with
        # the restored param_group rate while Adam's own step counter keeps counting.
        self.iteration = 0
        self.adam_steps = 0
        self._lr_now = self.lr
        if not bool(getattr(options.nerf, "use_viewdirs", True)):
            raise L.DexNerfError("Trainer: options.nerf.use_viewdirs is False, but the training kernels exist for "
                                 "FlexibleNeRFModel WITH view directions only")
        self.progs, self.specs, self.sizes = [], [], []
        for m in self.models:
            prog = m.program(self.ex, self.ed)
            spec = tensorcore.spec_for(m, prog) if tensorcore.trainable(m, prog) else None
            if spec is None:
                raise L.DexNerfError("Trainer needs tensor-core-capable FlexibleNeRFModels (view directions, "
                                     "hidden 128/256)")
            self.progs.append(prog)
            self.specs.append(spec)
            self.sizes.append(m.packed_params().numel())
        dev = self.models[0].packed_params().device
        total = sum(self.sizes)
        self.params = torch.empty(total, dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for m, n in zip(self.models, self.sizes):
            self.params[off:off + n].copy_(m.packed_params())
            self.views.append((off, n))
            off += n
        self.grads = torch.zeros_like(self.params)
        self._p2p = None             # data-parallel step over NVLink peer memory (see _setup_p2p)
        if self.world > 1 and os.environ.get("DEXNERF_P2P", "1") != "0":
            self._setup_p2p(dev, total)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.loss = torch.zeros(3, dtype=torch.float32, device=dev)     # total, coarse, fine
        self._chunks = {}
        self.keep_grads = False      # True: leave the step's gradients in `grads` (tests) instead of clearing them in Adam
        self.timing = None           # bench.py: (render.Events of the forward call, render.Events of the backward)
        self.blobs, self.blobs_t = [], []
        for spec in self.specs:
            self.blobs.append(torch.empty(L.lib().dexnerf_tc_packed_bytes(spec), dtype=torch.uint8, device=dev))
            self.blobs_t.append(torch.empty(L.lib().dexnerf_tc_packed_bwd_bytes(spec), dtype=torch.uint8, device=dev))
        self._repack()

    # -- NVLink peer memory ------------------------------------------------------------------
    def _setup_p2p(self, dev, total):
        """One process per GPU: two IPC-exportable gradient buffers (they ping-pong between steps) and a flag array per
        rank, every rank's mapped into every process.  With them a training step needs no NCCL call: the fused
        all-reduce + Adam kernel (csrc/p2p.cu) reads the peers' gradients directly.  Falls back to the NCCL all-reduce
        (self._p2p = None) when the ranks are not all P2P peers on one node or the handles cannot be opened."""
        import torch.distributed as dist
        rank = dist.get_rank(self.group)
        lib = L.lib()
        ok = total % 4 == 0 and self.world <= 8
        if ok:
            me = dev.index if dev.index is not None else torch.cuda.current_device()
            devs = [None] * self.world
            dist.all_gather_object(devs, (os.uname().nodename, me), group=self.group)
            ok = len({d[0] for d in devs}) == 1 and len({d[1] for d in devs}) == self.world
            ok = ok and all(d[1] == me or torch.cuda.can_device_access_peer(me, d[1]) for d in devs)
        flags_ok = [None] * self.world
        dist.all_gather_object(flags_ok, bool(ok), group=self.group)
        if not all(flags_ok):
            return
        own, handles = [], []
        for nbytes in (total * 4, total * 4, 256):          # gradient buffer 0, gradient buffer 1, flags
            ptr = C.c_void_p()
            L.check(lib.dexnerf_p2p_alloc(nbytes, C.byref(ptr)), "p2p_alloc")
            h = C.create_string_buffer(64)
            L.check(lib.dexnerf_p2p_export(ptr, h), "p2p_export")
            own.append(ptr.value)
            handles.append(h.raw)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, handles, group=self.group)
        table = []                                           # table[r] = (grads0, grads1, flags) of rank r, valid here
        opened = []
        try:
            for r in range(self.world):
                if r == rank:
                    table.append(tuple(own))
                    continue
                ptrs = []
                for h in everyone[r]:
                    q = C.c_void_p()
                    L.check(lib.dexnerf_p2p_open(C.create_string_buffer(h, 64), C.byref(q)), "p2p_open")
                    ptrs.append(q.value)
                    opened.append(q.value)
                table.append(tuple(ptrs))
            opened_ok = True
        except L.DexNerfError:
            opened_ok = False
        all_ok = [None] * self.world
        dist.all_gather_object(all_ok, opened_ok, group=self.group)
        if not all(all_ok):
            for q in opened:
                lib.dexnerf_p2p_close(q)
            for q in own:
                lib.dexnerf_p2p_free(q)
            return
        gbufs = [_wrap_device_memory(own[k], total, dev) for k in (0, 1)]
        arr = lambda k: (C.c_void_p * self.world)(*[table[r][k] for r in range(self.world)])
        self._p2p = dict(rank=rank, own=own, opened=opened, gbufs=gbufs, peer_grads=(arr(0), arr(1)), peer_flags=arr(2),
                         token=0)
        self.grads = gbufs[0]

    def close(self):
        """Unmap the peers' buffers and free this rank's (after a barrier: nobody may still be reading them)."""
        if self._p2p is not None:
            import torch.distributed as dist
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            p, self._p2p = self._p2p, None
            self.grads = p["gbufs"][0].clone()
            for q in p["opened"]:
                L.lib().dexnerf_p2p_close(q)
            dist.barrier(group=self.group)
            for q in p["own"]:
                L.lib().dexnerf_p2p_free(q)

    # -- parameter plumbing -------------------------------------------------------------------
    def _flat(self, buf, i):
        off, n = self.views[i]
        return buf[off:off + n]

    def _repack(self):
        """fp32 master parameters -> the bf16 operand images both directions stream (3 launches per net)."""
        for i, (spec, prog) in enumerate(zip(self.specs, self.progs)):
            p = self._flat(self.params, i)
            L.check(L.lib().dexnerf_tc_pack(spec, prog, L.ptr(p), L.ptr(self.blobs[i]), L.stream_ptr()), "tc_pack")
            L.launch_count += 1
            L.check(L.lib().dexnerf_tc_pack_bwd(spec, prog, L.ptr(p), L.ptr(self.blobs_t[i]), L.stream_ptr()),
                    "tc_pack_bwd")

    def sync_to_modules(self):
        """Flat master parameters -> nn.Module parameters (reference state-dict layout)."""
        with torch.no_grad():
            for i, m in enumerate(self.models):
                flat, prog = self._flat(self.params, i), self.progs[i]
                for k, (lin, *_r) in enumerate(m._layers()):
                    op = prog.ops[k]
                    fin, fout = lin.in_features, lin.out_features
                    lin.weight.copy_(flat[op.w_off:op.w_off + fin * fout].view(fin, fout).t())
                    lin.bias.copy_(flat[op.b_off:op.b_off + fout])

    def learning_rate(self):
        """The rate the next step uses (what the reference's optimizer.param_groups[0]["lr"] holds)."""
        return self._lr_now

    # -- checkpoints in the reference's format (train_dexnerf_rgb.py:442-457 / :167-174) --------
    def _param_slices(self):
        """(flat offset of W, flat offset of b, nn.Linear) for every parameter pair, in the order of
        list(model_coarse.parameters()) + list(model_fine.parameters()) - the optimizer's order
        (train_dexnerf_rgb.py:142-148)."""
        out = []
        for i, m in enumerate(self.models):
            base = self.views[i][0]
            by_lin = {id(lin): self.progs[i].ops[k] for k, (lin, *_r) in enumerate(m._layers())}
            for lin in (mod for mod in m.modules() if isinstance(mod, torch.nn.Linear)):
                op = by_lin.get(id(lin))
                out.append((None, None, lin) if op is None else (base + op.w_off, base + op.b_off, lin))
        return out

    def checkpoint_dict(self, loss=None, psnr=None):
        """{iter, model_coarse_state_dict, model_fine_state_dict, optimizer_state_dict, loss, psnr} as the
        reference writes it: the state dicts have the reference's keys, and the optimizer entry is a
        torch.optim.Adam state_dict over coarse + fine parameters, so the reference's script (or
        torch.optim.Adam.load_state_dict) resumes from it.  Conventions of the reference's writer: "iter" is the
        0-based loop index of the iteration just finished (k - 1 after k steps; :443), Adam's per-parameter
        "step" is the number of updates made, and the param_group "lr" is the rate set after that iteration
        (:283-289), i.e. the one the next iteration uses."""
        self.sync_to_modules()
        state, idx = {}, 0
        for w_off, b_off, lin in self._param_slices():
            fin, fout = lin.in_features, lin.out_features
            for off, shape, is_w in ((w_off, (fin, fout), True), (b_off, (fout,), False)):
                if off is None:          # a layer the forward never uses (none for FlexibleNeRFModel)
                    idx += 1
                    continue
                n = fin * fout if is_w else fout
                ea, es = self.exp_avg[off:off + n].view(shape), self.exp_avg_sq[off:off + n].view(shape)
                if is_w:
                    ea, es = ea.t(), es.t()
                state[idx] = {"step": torch.tensor(float(self.adam_steps)), "exp_avg": ea.contiguous().clone(),
                              "exp_avg_sq": es.contiguous().clone()}
                idx += 1
        group = {"lr": self._lr_now, "betas": self.betas, "eps": self.eps,
                 "weight_decay": 0, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
                 "differentiable": False, "fused": None, "decoupled_weight_decay": False, "params": list(range(idx))}
        if self.adam_steps == 0:
            state = {}
        return {"iter": max(self.iteration - 1, 0), "model_coarse_state_dict": self.models[0].state_dict(),
                "model_fine_state_dict": self.models[1].state_dict(),
                "optimizer_state_dict": {"state": state, "param_groups": [group]}, "loss": loss, "psnr": psnr}

    def load_checkpoint_dict(self, ckpt):
        """Resume from a checkpoint written by the reference script or by checkpoint_dict()."""
        self.models[0].load_state_dict(ckpt["model_coarse_state_dict"])
        if ckpt.get("model_fine_state_dict"):
            self.models[1].load_state_dict(ckpt["model_fine_state_dict"])
        off = 0
        for m, n in zip(self.models, self.sizes):
            m.__dict__.pop("_packed_cache", None)
            self.params[off:off + n].copy_(m.packed_params())
            off += n
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        state = (ckpt.get("optimizer_state_dict") or {}).get("state", {})
        idx = 0
        for w_off, b_off, lin in self._param_slices():
            fin, fout = lin.in_features, lin.out_features
            for off_, is_w in ((w_off, True), (b_off, False)):
                st = state.get(idx)
                idx += 1
                if st is None or off_ is None:
                    continue
                n = fin * fout if is_w else fout
                ea, es = st["exp_avg"].to(self.params.device), st["exp_avg_sq"].to(self.params.device)
                if is_w:
                    ea, es = ea.t().contiguous(), es.t().contiguous()
                self.exp_avg[off_:off_ + n].copy_(ea.reshape(-1))
                self.exp_avg_sq[off_:off_ + n].copy_(es.reshape(-1))
        # resume as the reference script does (train_dexnerf_rgb.py:167-178): the loop restarts AT index `iter`,
        # the optimizer keeps its own step counter and the rate of its restored param_group
        self.iteration = int(ckpt.get("iter", 0))
        steps = [float(st["step"]) for st in state.values() if "step" in st]
        self.adam_steps = int(max(steps)) if steps else 0
        groups = (ckpt.get("optimizer_state_dict") or {}).get("param_groups") or []
        self._lr_now = float(groups[0]["lr"]) if groups and "lr" in groups[0] else self.lr
        self._repack()

    # -- one iteration ------------------------------------------------------------------------
    def _chunk_buffers(self, n):
        """Device buffers of one ray chunk of `n` rays (cached per size; a training loop uses one size): the
        workspace of the fused render call, the two tapes, the rgb predictions, their gradients and the
        d(radiance field) scratch."""
        from . import render
        buf = self._chunks.get(n)
        if buf is None:
            opt = self.cfg.nerf.train
            Nc, Nf = int(opt.num_coarse), int(opt.num_fine)
            dev = self.params.device
            self._chunks.clear()                      # a new size replaces the old buffers (tapes are GBs)
            tapes = []
            for spec, S in zip(self.specs, (Nc, Nc + Nf)):
                nbytes = L.lib().dexnerf_tc_tape_bytes(spec, n * S)
                if nbytes < 0:
                    raise L.DexNerfError("tc_tape_bytes: " + L.lib().dexnerf_last_error().decode())
                tapes.append(torch.empty(nbytes, dtype=torch.uint8, device=dev))
            buf = dict(ws=torch.empty(render.workspace_bytes(n, Nc, Nf), dtype=torch.uint8, device=dev), tapes=tapes,
                       rgb=torch.empty((2, n, 3), dtype=torch.float32, device=dev),
                       g_rgb=torch.empty((2, n, 3), dtype=torch.float32, device=dev),
                       d_rf=torch.empty((n, Nc + Nf, 4), dtype=torch.float32, device=dev))
            self._chunks[n] = buf
        return buf

    def _render_params(self, n, buf, ro, rd, rng, height, width, focal_length):
        from . import render
        opt = self.cfg.nerf.train
        p = L.RenderParams()
        render.fill_common(p, opt, self.cfg, True, height, width, focal_length, None, 0)
        p.n = n
        p.ro, p.rd = ro.data_ptr(), rd.data_ptr()
        keep = []
        for i, m in enumerate(self.models):
            ref, k = render.model_ref(m, self.ex, self.ed, "bf16",
                                      blobs=(self.blobs[i], self.blobs_t[i], self._flat(self.params, i)))
            keep.append(k)
            if i == 0:
                p.coarse = ref
            else:
                p.fine = ref
        p.t_rand, p.u = L.ptr(rng.get("t_rand")).value, L.ptr(rng.get("u")).value
        p.noise_coarse, p.noise_fine = L.ptr(rng.get("noise_coarse")).value, L.ptr(rng.get("noise_fine")).value
        p.offset = render.next_philox_offset()
        p.tape_coarse, p.tape_fine = buf["tapes"][0].data_ptr(), buf["tapes"][1].data_ptr()
        p.workspace, p.workspace_bytes = buf["ws"].data_ptr(), buf["ws"].numel()
        p.rgb_coarse, p.rgb_fine = buf["rgb"][0].data_ptr(), buf["rgb"][1].data_ptr()
        return p, keep

    def step(self, ray_origins, ray_directions, target, rng=None, height=None, width=None, focal_length=None):
        """One iteration on pre-selected rays (n,3) / targets (n,3).  Returns the loss tensor
        [total, coarse, fine] (device, overwritten by the next step).  `rng` replays the four draws; without it
        they come from the Philox generator of the setup launch.
        Batches larger than `cfg.nerf.train.chunksize` rays are processed in chunks (the tape of a chunk
        is 10 KB per sample) whose gradients accumulate in the flat buffer before the single Adam step.
        With `cfg.dataset.no_ndc: False` (the LLFF configs) the rays are warped by ndc_rays exactly as
        run_one_iter_of_nerf does (train_utils.py:238-242), which needs `height`, `width` and `focal_length`."""
        rng = {k: L.dev_f32(v, k) for k, v in (rng or {}).items() if v is not None}
        opt = self.cfg.nerf.train
        if self.cfg.dataset.no_ndc is False and (height is None or width is None or focal_length is None):
            raise L.DexNerfError("Trainer.step: cfg.dataset.no_ndc is False - pass height, width and focal_length "
                                 "(ndc_rays needs them, train_utils.py:238-242)")
        ro = L.dev_f32(ray_origins.reshape(-1, 3), "ray_origins")
        rd = L.dev_f32(ray_directions.reshape(-1, 3), "ray_directions")
        tgt = L.dev_f32(target[..., :3].reshape(-1, 3), "target")
        n_total = ro.shape[0]
        chunk = int(getattr(opt, "chunksize", n_total) or n_total)
        self.loss.zero_()
        if self.keep_grads and self._p2p is None:
            self.grads.zero_()        # otherwise the previous Adam launch has already cleared them
        starts = list(range(0, n_total, chunk))
        pending = []
        for start in starts:
            sl = slice(start, min(start + chunk, n_total))
            sub = {k: v[sl] for k, v in rng.items()}
            # the fine network's backward runs first: on the last chunk its gradients are final as soon as its
            # weight-gradient GEMM is enqueued, so their all-reduce overlaps the coarse network's backward
            after_fine = None
            if self.world > 1 and self._p2p is None and start == starts[-1]:
                after_fine = lambda: pending.append(self._allreduce_async(self._flat(self.grads, 1)))
            self._accumulate(ro[sl], rd[sl], tgt[sl], sub, n_total, after_fine, height, width, focal_length)
        if self.world > 1 and self._p2p is None:
            pending.append(self._allreduce_async(self._flat(self.grads, 0)))
            for work in pending:
                work.wait()              # stream-level wait: the Adam launch below is ordered after both
        # Adam with the script's schedule: iteration i steps with the rate set after iteration i-1 (:283-289)
        lr = self._lr_now
        self.adam_steps += 1
        if self._p2p is not None:
            # ONE kernel: barrier over NVLink flags, sum of every rank's gradients (P2P loads, rank order), Adam,
            # and the other gradient buffer cleared for the next step (csrc/p2p.cu)
            p = self._p2p
            cur = p["token"] & 1
            p["token"] += 1
            nxt = p["gbufs"][cur ^ 1]
            L.check(L.lib().dexnerf_adam_step_allreduce(
                L.ptr(self.params), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), L.ptr(nxt), self.params.numel(),
                p["peer_grads"][cur], p["peer_flags"], p["rank"], self.world, p["token"], lr, self.betas[0],
                self.betas[1], self.eps, self.adam_steps, 1.0 / self.world, L.stream_ptr()), "adam_step_allreduce")
            if self.keep_grads:
                self.last_grads = p["gbufs"][cur]        # this rank's own gradients of the step (not the mean)
            self.grads = nxt
        else:
            adam = L.lib().dexnerf_adam_step if self.keep_grads else L.lib().dexnerf_adam_step_zero_grad
            L.check(adam(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                         self.params.numel(), lr, self.betas[0], self.betas[1], self.eps, self.adam_steps,
                         1.0 / self.world, L.stream_ptr()), "adam_step")
        self._lr_now = learning_rate(self.lr, self.iteration, self.lr_decay, self.lr_decay_factor)
        self.iteration += 1
        self._repack()
        return self.loss

    def _allreduce_async(self, buf):
        import torch.distributed as dist
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _accumulate(self, ro, rd, tgt, rng, n_total, after_fine=None, height=None, width=None, focal_length=None):
        """Forward + backward of one ray chunk: ONE fused forward call (6 launches: setup incl. the Philox draws,
        two queries with tape, two compositings, resampling), one loss launch, the backward of the fine
        and then the coarse network (3 launches each: compositing backward, activation-gradient chain,
        weight-gradient GEMM); loss terms and gradients accumulate."""
        from . import render
        ro, rd, tgt = ro.contiguous(), rd.contiguous(), tgt.contiguous()
        n = ro.shape[0]
        buf = self._chunk_buffers(n)
        p, keep = self._render_params(n, buf, ro, rd, rng, height, width, focal_length)
        lib, stream = L.lib(), L.stream_ptr()
        ev = self.timing
        p.events = ev[0].pointer() if ev else None
        L.check(lib.dexnerf_render_fused_fwd(C.byref(p), stream), "render_fused_fwd")
        L.launch_count += 5 + p.ndc
        # loss = mse(rgb_coarse, target) + mse(rgb_fine, target)  (train_dexnerf_rgb.py:264-277)
        g = buf["g_rgb"]
        L.check(lib.dexnerf_mse_loss_pair(L.ptr(buf["rgb"][0]), L.ptr(buf["rgb"][1]), L.ptr(tgt), 3 * n, 3 * n_total,
                                          L.ptr(g[0]), L.ptr(g[1]), L.ptr(self.loss), stream), "mse_loss_pair")
        p.events = ev[1].pointer() if ev else None
        gc, gf = self._flat(self.grads, 0), self._flat(self.grads, 1)
        # Without a hook between the two networks ONE call runs both backwards, the coarse network's chain on a second
        # stream next to the fine network's weight-gradient GEMM (render.cu: DEXNERF_BWD_SPLIT).
        for which in ((3,) if after_fine is None else (1, 2)):
            L.check(lib.dexnerf_render_fused_bwd(C.byref(p), L.ptr(g[0]), L.ptr(g[1]), L.ptr(buf["d_rf"]), L.ptr(gc),
                                                 L.ptr(gf), which, stream), "render_fused_bwd")
            L.launch_count += 2 * (2 if which == 3 else 1)   # compositing backward, activation-gradient chain, weight-gradient GEMM
            if which == 1 and after_fine is not None:
                after_fine()
        del keep
