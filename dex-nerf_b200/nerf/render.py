"""Host side of the fused render driver (csrc/render.cu, `dexnerf_render_fused_fwd`): one C-ABI call per ray
chunk replaces the reference's predict_and_render_radiance (nerf/train_utils.py:92-202) - ray packing, view
directions, near / far columns, stratified depths, the four random draws, both network passes, resampling and
both compositings - with no torch kernel in between.  This module only fills the parameter struct: outputs and
the workspace are torch allocations (plumbing), everything numeric happens behind include/dexnerf.h."""
import ctypes as C

import torch

from . import _lib as L
from . import tensorcore

_workspaces = {}       # device index -> uint8 tensor, grown on demand and reused by every inference call
_philox_offset = [0]   # advances once per fused call that draws random numbers (like torch's generator offset)

LAUNCH_NAMES = ("ray_setup", "ndc_rays", "mlp_coarse", "composite_coarse", "resample_merge", "mlp_fine",
                "composite_fine")
# dexnerf_render_fused_bwd's event slots: slots 1 / 4 are the activation-gradient chain and 2 / 5 the weight-gradient
# GEMM (under DEXNERF_BWD=fused / shared slot 1 / 4 is the ONE MLP backward launch and 2 / 5 stay empty)
BWD_LAUNCH_NAMES = ("composite_bwd_fine", "mlp_bwd_fine", "mlp_bwd_dw_fine", "composite_bwd_coarse",
                    "mlp_bwd_coarse", "mlp_bwd_dw_coarse")


def workspace_bytes(n, Nc, Nf):
    nbytes = L.lib().dexnerf_render_workspace_bytes(int(n), int(Nc), int(Nf))
    if nbytes < 0:
        raise L.DexNerfError("render_workspace_bytes: " + L.lib().dexnerf_last_error().decode())
    return int(nbytes)


def workspace_layout(n, Nc, Nf):
    """Byte offsets of the intermediates inside a workspace (tests / the training paths read them back)."""
    out = (C.c_int64 * L.RENDER_WS_SLOTS)()
    rc = L.lib().dexnerf_render_workspace_layout(int(n), int(Nc), int(Nf), out)
    if rc != 0:
        raise L.DexNerfError("render_workspace_layout: " + L.lib().dexnerf_last_error().decode())
    names = ("total", "ro", "rd", "viewdirs", "z_coarse", "rf_coarse", "weights_coarse", "z_fine", "rf_fine",
             "t_rand", "u", "noise_coarse", "noise_fine", "ro_raw", "rd_raw")
    return {k: int(out[i]) for i, k in enumerate(names)}


def shared_workspace(nbytes, device):
    """The per-device scratch buffer of the inference path (torch's caching allocator hands out 512-byte aligned
    blocks, the C side asks for 256)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        _workspaces.pop(key, None)
        ws = None
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def ws_view(ws, layout, name, shape):
    """A float32 view of one workspace slot."""
    numel = 1
    for s in shape:
        numel *= int(s)
    off = layout[name]
    return ws[off:off + 4 * numel].view(torch.float32).view(*shape)


def model_ref(model, embed_fn, embeddirs_fn, precision, blobs=None):
    """dexnerf_model_ref of `model` plus the objects that must stay alive while the call runs.  `blobs`
    (packed, packed_t, params) overrides the module's own cached weight images (nerf.Trainer owns flat copies)."""
    prog = model.program(embed_fn, embeddirs_fn if model.dim_dir else None)
    ref = L.ModelRef()
    keep = [prog]
    ref.prog = C.pointer(prog)
    spec = tensorcore.spec_for(model, prog) if precision == "bf16" else None
    if blobs is not None:
        packed, packed_t, params = blobs
        keep.append(spec)
        ref.spec = C.pointer(spec)
        ref.packed, ref.packed_t, ref.params = packed.data_ptr(), packed_t.data_ptr(), params.data_ptr()
        keep += [packed, packed_t, params]
        return ref, keep
    params = model.packed_params()
    ref.params = params.data_ptr()
    keep.append(params)
    if spec is not None:
        blob = tensorcore.packed_weights(model, prog, spec)
        ref.spec = C.pointer(spec)
        ref.packed = blob.data_ptr()
        keep += [spec, blob]
    return ref, keep


def next_philox_offset():
    _philox_offset[0] += 1
    return _philox_offset[0]


class Events:
    """2 x RENDER_LAUNCHES CUDA events owned by the library (bench.py's per-kernel breakdown)."""

    def __init__(self):
        lib = L.lib()
        self.arr = (C.c_void_p * (2 * L.RENDER_LAUNCHES))()
        for i in range(2 * L.RENDER_LAUNCHES):
            self.arr[i] = lib.dexnerf_event_create()
            if not self.arr[i]:
                raise L.DexNerfError("event_create failed")

    def pointer(self):
        return C.cast(self.arr, C.POINTER(C.c_void_p))

    def elapsed_ms(self, names=LAUNCH_NAMES):
        """{launch name: ms} for the launches that ran in the last call (after a synchronize)."""
        lib, out = L.lib(), {}
        for k, name in enumerate(names):
            ms = lib.dexnerf_event_elapsed_ms(self.arr[2 * k], self.arr[2 * k + 1])
            if ms >= 0.0:
                out[name] = float(ms)
        return out

    def __del__(self):
        try:
            lib = L.lib()
            for e in self.arr:
                lib.dexnerf_event_destroy(e)
        except Exception:
            pass


event_hook = None      # bench.py: callable(n) -> Events or None, consulted once per fused call


def fill_common(p, opt, options, use_viewdirs, height, width, focal_length, thr, T):
    p.use_viewdirs = int(bool(use_viewdirs))
    p.ndc = int(options.dataset.no_ndc is False)
    p.H, p.W = int(height or 0), int(width or 0)
    p.focal = float(focal_length) if focal_length is not None else 0.0
    p.near, p.far = float(options.dataset.near), float(options.dataset.far)
    p.Nc, p.Nf = int(opt.num_coarse), int(opt.num_fine)
    p.lindisp = int(bool(opt.lindisp))
    p.perturb = int(bool(opt.perturb))
    p.white_background = int(bool(opt.white_background))
    p.noise_std = float(opt.radiance_field_noise_std)
    p.thresholds, p.T = (thr.data_ptr() if T else None), int(T)
    p.seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF


def _ptr_at(t, row, row_floats):
    return None if t is None else t.data_ptr() + 4 * int(row) * int(row_floats)


def render_rays(height, width, focal_length, model_coarse, model_fine, ro, rd, options, mode, embed_fn, embeddirs_fn,
                thr, T, rng, precision, camera=None, out=None):
    """All chunks of one run_one_iter_of_nerf call through the fused entry.  ro, rd: contiguous (n,3) CUDA fp32 -
    or None with camera = (T_w2c, K, row0, rows) (ray generation inside the setup launch).
    Returns [rgb_c, depth_c, acc_c, rgb_f, depth_f, acc_f, dex (T,n) or None] as flat tensors."""
    opt = getattr(options.nerf, mode)
    dev = ro.device if ro is not None else camera[0].device
    if camera is not None:
        T_w2c, K, row0, rows = camera
        n_total = int(rows) * int(width)
    else:
        n_total = ro.shape[0]
    Nc, Nf = int(opt.num_coarse), int(opt.num_fine)
    outs = [torch.empty((n_total, 3), dtype=torch.float32, device=dev), torch.empty((n_total,), dtype=torch.float32, device=dev),
            torch.empty((n_total,), dtype=torch.float32, device=dev), torch.empty((n_total, 3), dtype=torch.float32, device=dev),
            torch.empty((n_total,), dtype=torch.float32, device=dev), torch.empty((n_total,), dtype=torch.float32, device=dev),
            torch.empty((T, n_total), dtype=torch.float32, device=dev) if T else None]
    dex_stride = n_total
    if out is not None:
        # fine-pass planes written in place, e.g. straight into another GPU's frame (nerf.SharedFrame.outputs):
        # [rgb (n,3), depth (n,), acc (n,), first Dex plane (n,) or None, floats between Dex planes]
        for k, t in enumerate(out[:4]):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.shape[0] != n_total):
                raise ValueError("render_rays: out[%d] must be a contiguous float32 tensor of %d rays" % (k, n_total))
        outs[3], outs[4], outs[5] = out[0], out[1], out[2]
        if T:
            outs[6], dex_stride = out[3], int(out[4])
    if n_total == 0:
        return outs
    use_viewdirs = bool(options.nerf.use_viewdirs)
    ref_c, keep_c = model_ref(model_coarse, embed_fn, embeddirs_fn, precision)
    ref_f, keep_f = model_ref(model_fine, embed_fn, embeddirs_fn, precision)
    chunk = int(opt.chunksize)
    if camera is not None:
        chunk = max(int(width), chunk // int(width) * int(width))      # whole image rows per chunk
    ws = shared_workspace(workspace_bytes(min(chunk, n_total), Nc, Nf), dev)
    rng = {k: L.dev_f32(v, k) for k, v in (rng or {}).items() if v is not None}
    draws = (opt.perturb and ("t_rand" not in rng or "u" not in rng)) or (
        float(opt.radiance_field_noise_std) > 0.0 and ("noise_coarse" not in rng or "noise_fine" not in rng))
    lib, stream = L.lib(), L.stream_ptr()
    p = L.RenderParams()
    fill_common(p, opt, options, use_viewdirs, height, width, focal_length, thr, T)
    p.coarse, p.fine = ref_c, ref_f
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    p.dex_stride = dex_stride
    for start in range(0, n_total, chunk):
        n = min(chunk, n_total - start)
        p.n = n
        if camera is not None:
            p.ro = p.rd = None
            p.T_w2c, p.K = T_w2c.data_ptr(), K.data_ptr()
            p.row0, p.rows = int(row0) + start // int(width), n // int(width)
        else:
            p.ro, p.rd = _ptr_at(ro, start, 3), _ptr_at(rd, start, 3)
        p.t_rand = _ptr_at(rng.get("t_rand"), start, Nc)
        p.u = _ptr_at(rng.get("u"), start, Nf)
        p.noise_coarse = _ptr_at(rng.get("noise_coarse"), start, Nc)
        p.noise_fine = _ptr_at(rng.get("noise_fine"), start, Nc + Nf)
        p.offset = next_philox_offset() if draws else 0
        p.rgb_coarse, p.depth_coarse, p.acc_coarse = _ptr_at(outs[0], start, 3), _ptr_at(outs[1], start, 1), _ptr_at(outs[2], start, 1)
        p.rgb_fine, p.depth_fine, p.acc_fine = _ptr_at(outs[3], start, 3), _ptr_at(outs[4], start, 1), _ptr_at(outs[5], start, 1)
        p.dex_fine = _ptr_at(outs[6], start, 1)
        ev = event_hook(n) if event_hook is not None else None
        p.events = ev.pointer() if ev is not None else None
        L.check(lib.dexnerf_render_fused_fwd(C.byref(p), stream), "render_fused_fwd")
        L.launch_count += 5 + p.ndc          # the call is 6 launches (7 with ndc_rays); check() counted one
    del keep_c, keep_f
    return outs
