"""ctypes binding of libdexnerf.so (include/dexnerf.h) - the only bridge between the Python
namespace and the CUDA kernels.  There is no fallback: if the library is missing or a tensor is
not a contiguous fp32 CUDA tensor, the call raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DEXNERF_LIB: an experiment build of the SAME library (dex-nerf_b200/build.py --define ... --out ...), tools/ only
LIB_PATH = os.environ.get("DEXNERF_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libdexnerf.so")

MAX_OPS = 16
ENC_XYZ, ENC_DIR, BUF_A, BUF_B, OUT_RGB, OUT_SIGMA, OUT_ALL, NONE = 0, 1, 2, 3, 4, 5, 6, -1


class Op(C.Structure):
    _fields_ = [("src0", C.c_int32), ("src0_dim", C.c_int32), ("src1", C.c_int32), ("src1_dim", C.c_int32),
                ("dst", C.c_int32), ("out_dim", C.c_int32), ("relu", C.c_int32), ("pad_", C.c_int32),
                ("w_off", C.c_int64), ("b_off", C.c_int64)]


class Program(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("dim_xyz", C.c_int32), ("dim_dir", C.c_int32), ("max_width", C.c_int32),
                ("Lx", C.c_int32), ("Ld", C.c_int32), ("include_xyz", C.c_int32), ("include_dir", C.c_int32),
                ("log_xyz", C.c_int32), ("log_dir", C.c_int32), ("pad_", C.c_int32 * 2),
                ("ops", Op * MAX_OPS)]


class FlexibleSpec(C.Structure):
    _fields_ = [("hidden", C.c_int32), ("n_trunk", C.c_int32), ("skip_every", C.c_int32),
                ("dim_xyz", C.c_int32), ("dim_dir", C.c_int32), ("Lx", C.c_int32), ("Ld", C.c_int32),
                ("include_xyz", C.c_int32), ("include_dir", C.c_int32), ("log_xyz", C.c_int32),
                ("log_dir", C.c_int32), ("arch", C.c_int32)]


class ModelRef(C.Structure):
    _fields_ = [("prog", C.POINTER(Program)), ("params", C.c_void_p), ("spec", C.POINTER(FlexibleSpec)),
                ("packed", C.c_void_p), ("packed_t", C.c_void_p)]


class RenderParams(C.Structure):
    """dexnerf_render_params (include/dexnerf.h): one ray chunk of the fused render driver."""
    _fields_ = [("n", C.c_int64), ("ro", C.c_void_p), ("rd", C.c_void_p), ("T_w2c", C.c_void_p), ("K", C.c_void_p),
                ("H", C.c_int32), ("W", C.c_int32), ("row0", C.c_int32), ("rows", C.c_int32),
                ("use_viewdirs", C.c_int32), ("ndc", C.c_int32),
                ("focal", C.c_float), ("near", C.c_float), ("far", C.c_float),
                ("Nc", C.c_int32), ("Nf", C.c_int32), ("lindisp", C.c_int32), ("perturb", C.c_int32),
                ("white_background", C.c_int32), ("noise_std", C.c_float),
                ("thresholds", C.c_void_p), ("T", C.c_int32), ("pad0_", C.c_int32),
                ("t_rand", C.c_void_p), ("u", C.c_void_p), ("noise_coarse", C.c_void_p), ("noise_fine", C.c_void_p),
                ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("coarse", ModelRef), ("fine", ModelRef),
                ("tape_coarse", C.c_void_p), ("tape_fine", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("rgb_coarse", C.c_void_p), ("depth_coarse", C.c_void_p), ("acc_coarse", C.c_void_p),
                ("rgb_fine", C.c_void_p), ("depth_fine", C.c_void_p), ("acc_fine", C.c_void_p),
                ("dex_fine", C.c_void_p), ("dex_stride", C.c_int64),
                ("events", C.POINTER(C.c_void_p))]


RENDER_LAUNCHES = 7
RENDER_WS_SLOTS = 16

_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    "dexnerf_abi_version": (C.c_int, []),
    "dexnerf_last_error": (C.c_char_p, []),
    "dexnerf_ray_bundle": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "dexnerf_ndc_rays": (C.c_int, [_P, _P, _L, _I, _I, _F, _F, _P, _P, _P]),
    "dexnerf_positional_encoding": (C.c_int, [_P, _L, _I, _I, _I, _P, _P]),
    "dexnerf_stratified_z": (C.c_int, [_L, _I, _F, _F, _P, _P, _I, _P, _P, _P]),
    "dexnerf_cumprod_exclusive": (C.c_int, [_P, _L, _I, _P, _P]),
    "dexnerf_volume_render": (C.c_int, [_P, _P, _P, _P, _L, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "dexnerf_sample_pdf": (C.c_int, [_P, _P, _L, _I, _I, _P, _P, _P, _P]),
    "dexnerf_resample_merge": (C.c_int, [_P, _P, _L, _I, _I, _P, _P, _P]),
    "dexnerf_mlp_forward": (C.c_int, [C.POINTER(Program), _P, _P, _L, _P, _P]),
    "dexnerf_mlp_query": (C.c_int, [C.POINTER(Program), _P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "dexnerf_tc_packed_bytes": (C.c_int64, [C.POINTER(FlexibleSpec)]),
    "dexnerf_tc_pack": (C.c_int, [C.POINTER(FlexibleSpec), C.POINTER(Program), _P, _P, _P]),
    "dexnerf_tc_query": (C.c_int, [C.POINTER(FlexibleSpec), _P, _P, _P, _P, _P, _L, _I, _P, _P, _I, _I, _P]),
    "dexnerf_volume_render_backward": (C.c_int, [_P, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P]),
    "dexnerf_mse_loss_grad": (C.c_int, [_P, _P, _L, _L, _P, _P, _P]),
    "dexnerf_mse_loss_pair": (C.c_int, [_P, _P, _P, _L, _L, _P, _P, _P, _P]),
    "dexnerf_adam_step_zero_grad": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _P]),
    "dexnerf_pack_params": (C.c_int, [C.POINTER(Program), _P, _P, _P]),
    "dexnerf_adam_step": (C.c_int, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _P]),
    "dexnerf_p2p_alloc": (C.c_int, [_L, C.POINTER(C.c_void_p)]),
    "dexnerf_p2p_free": (C.c_int, [_P]),
    "dexnerf_p2p_export": (C.c_int, [_P, _P]),
    "dexnerf_p2p_open": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "dexnerf_p2p_close": (C.c_int, [_P]),
    "dexnerf_adam_step_allreduce": (C.c_int, [_P, _P, _P, _P, _L, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                              C.c_int, C.c_uint32, _F, _F, _F, _F, _L, _F, _P]),
    "dexnerf_depth_error_metrics": (C.c_int, [_P, _P, _P, _L, _I, _P, _P, _P, _P]),
    "dexnerf_depth_error_image": (C.c_int, [_P, _P, _P, _I, _I, _F, _P, _P]),
    "dexnerf_render_workspace_bytes": (C.c_int64, [_L, _I, _I]),
    "dexnerf_render_workspace_layout": (C.c_int, [_L, _I, _I, C.POINTER(C.c_int64)]),
    "dexnerf_ray_setup": (C.c_int, [C.POINTER(RenderParams), _P]),
    "dexnerf_render_fused_fwd": (C.c_int, [C.POINTER(RenderParams), _P]),
    "dexnerf_render_fused_bwd": (C.c_int, [C.POINTER(RenderParams), _P, _P, _P, _P, _P, _I, _P]),
    "dexnerf_event_create": (C.c_void_p, []),
    "dexnerf_event_destroy": (None, [_P]),
    "dexnerf_event_elapsed_ms": (C.c_float, [_P, _P]),
    "dexnerf_tc_tape_bytes": (C.c_int64, [C.POINTER(FlexibleSpec), _L]),
    "dexnerf_tc_tape_layout": (C.c_int, [C.POINTER(FlexibleSpec), _L, C.POINTER(C.c_int64)]),
    "dexnerf_tc_query_train": (C.c_int, [C.POINTER(FlexibleSpec), _P, _P, _P, _P, _P, _L, _I, _P, _P, _P]),
    "dexnerf_tc_packed_bwd_bytes": (C.c_int64, [C.POINTER(FlexibleSpec)]),
    "dexnerf_tc_pack_bwd": (C.c_int, [C.POINTER(FlexibleSpec), C.POINTER(Program), _P, _P, _P]),
    "dexnerf_tc_backward": (C.c_int, [C.POINTER(FlexibleSpec), C.POINTER(Program), _P, _P, _P, _P, _L, _I, _P, _I, _I,
                                      _P]),
}

_lib = None


class DexNerfError(RuntimeError):
    pass


def lib():
    """Load libdexnerf.so once.  Raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DexNerfError(
                "libdexnerf.so not found at %s - build it with `python dex-nerf_b200/build.py` "
                "(there is no CPU or PyTorch fallback)" % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


launch_count = 0   # C-ABI calls made so far; every call is exactly one kernel launch


def check(rc, what):
    global launch_count
    launch_count += 1
    if rc != 0:
        raise DexNerfError("%s failed (%d): %s" % (what, rc, lib().dexnerf_last_error().decode()))


event_log = None     # bench.py sets this to a list: (kernel name, start event, end event, n, S, algorithmic bytes)


class timed:
    """CUDA-event bracket around one C-ABI call, active only while bench.py collects a breakdown."""

    def __init__(self, name, n, S, nbytes=0):
        self.rec = (name, n, S, nbytes)

    def __enter__(self):
        if event_log is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if event_log is not None:
            self.e1.record()
            name, n, S, nbytes = self.rec
            event_log.append((name, self.e0, self.e1, n, S, nbytes))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev_f32(t, name, allow_none=False):
    """Validate a tensor argument: contiguous fp32 on CUDA.  No silent copies to/from the host."""
    if t is None:
        if allow_none:
            return None
        raise ValueError("%s is required" % name)
    if not isinstance(t, torch.Tensor):
        raise ValueError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (this build has no CPU path)" % name)
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32, got %s" % (name, t.dtype))
    if t.device.index != torch.cuda.current_device():
        # the C ABI launches on the CURRENT device and stream; a tensor of another GPU would be read through
        # an invalid address there
        raise ValueError("%s lives on %s but the current CUDA device is cuda:%d (use torch.cuda.set_device / "
                         "torch.cuda.device)" % (name, t.device, torch.cuda.current_device()))
    return t.contiguous()


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())
