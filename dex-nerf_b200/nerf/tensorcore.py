"""Host side of the tcgen05 tensor-core MLP path (dex-nerf_b200/csrc/mlp_tc.cu): decides whether a
model can run on it, packs its weights into bf16 UMMA shared-memory images (cached until a
parameter changes) and launches the fused encode + MLP kernel."""
import torch

from . import _lib as L
from .models import FlexibleNeRFModel, PaperNeRFModel


def spec_for(model, prog):
    """FlexibleSpec for a model the kernel supports - FlexibleNeRFModel (hidden 128 / 256; with view directions:
    inference and training, without: arch = 2, inference only) or PaperNeRFModel (arch = 1, inference only) -
    else None."""
    if not isinstance(model, (FlexibleNeRFModel, PaperNeRFModel)):
        return None
    no_dirs = isinstance(model, FlexibleNeRFModel) and not model.use_viewdirs
    if not model.use_viewdirs and not no_dirs:
        return None
    if not (1 <= model.dim_xyz <= 64 and (model.dim_dir == 0 if no_dirs else 1 <= model.dim_dir <= 32)):
        return None
    s = L.FlexibleSpec()
    if isinstance(model, PaperNeRFModel):
        s.hidden, s.n_trunk, s.skip_every, s.arch = 256, 7, 4, 1
    else:
        H = model.hidden_size
        n_trunk = len(model.layers_xyz)
        if H not in (128, 256) or n_trunk < 1 or n_trunk + 3 > 16:
            return None
        s.hidden, s.n_trunk, s.skip_every, s.arch = H, n_trunk, int(model.skip_connect_every), 2 if no_dirs else 0
    s.dim_xyz, s.dim_dir = int(model.dim_xyz), int(model.dim_dir)
    s.Lx, s.Ld = prog.Lx, prog.Ld
    s.include_xyz, s.include_dir, s.log_xyz, s.log_dir = prog.include_xyz, prog.include_dir, prog.log_xyz, prog.log_dir
    return s


def supported(model, prog):
    return spec_for(model, prog) is not None


def trainable(model, prog):
    """The tape / backward kernels exist for the FlexibleNeRFModel layer table only."""
    return isinstance(model, FlexibleNeRFModel) and model.use_viewdirs and supported(model, prog)


def packed_weights(model, prog, spec):
    params = model.packed_params()
    cache = model.__dict__.get("_tc_cache")
    key = (params.data_ptr(), model.__dict__["_packed_cache"][0])
    if cache is None or cache[0] != key:
        nbytes = L.lib().dexnerf_tc_packed_bytes(spec)
        if nbytes <= 0:
            raise L.DexNerfError("tc_packed_bytes: " + L.lib().dexnerf_last_error().decode())
        blob = torch.empty(nbytes, dtype=torch.uint8, device=params.device)
        L.check(L.lib().dexnerf_tc_pack(spec, prog, L.ptr(params), L.ptr(blob), L.stream_ptr()), "tc_pack")
        cache = (key, blob)
        model.__dict__["_tc_cache"] = cache
    return cache[1]


def query(model, prog, ro, rd, viewdirs, z, rf, dbg=None, dbg_layer=-1, dbg_pass=0):
    spec = spec_for(model, prog)
    if spec is None:
        raise L.DexNerfError("model is not supported by the tensor-core path")
    blob = packed_weights(model, prog, spec)
    n, S = z.shape
    L.check(L.lib().dexnerf_tc_query(spec, L.ptr(blob), L.ptr(ro), L.ptr(rd), L.ptr(viewdirs), L.ptr(z), n, S,
                                     L.ptr(rf), L.ptr(dbg), int(dbg_layer), int(dbg_pass), L.stream_ptr()),
            "tc_query")
    return rf
