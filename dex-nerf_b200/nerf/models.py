"""Drop-in for the reference's nerf/models.py.  Every class is a torch.nn.Module with the same
constructor arguments, parameter names, shapes and creation order (hence identical default
initialisation under a seed and loadable pretrained/*.ckpt state dicts).  `forward` does not call
torch.nn.functional.linear: each model lowers to a layer program (include/dexnerf.h,
dexnerf_mlp_program) that the CUDA kernels interpret.

Two forwards differ from the reference ON PURPOSE because the reference's raise for the 8-layer
configurations; the repairs are the ones the oracle documents:
  * FlexibleNeRFModel: skip condition `i % skip_connect_every == 0 and i > 0` (what __init__
    builds, models.py:210) instead of the attribute error at models.py:243;
  * PaperNeRFModel: the xyz trunk starts from xyz only (models.py:165-169 feed xyz+dirs).
"""
import torch

from . import _lib as L

_Linear = torch.nn.Linear


def _align4(n):
    """Blocks of the flat parameter buffer start on multiples of 4 floats (vector loads / reductions)."""
    return (n + 3) & ~3


class _ProgramModule(torch.nn.Module):
    """Shared lowering/packing logic.  Subclasses implement `_layers()` -> list of
    (linear_module, src0, src1, dst, relu) with buffer ids from _lib."""

    dim_xyz = 0
    dim_dir = 0

    def _layers(self):
        raise NotImplementedError

    # -- lowering ------------------------------------------------------------------------
    def program(self, embed_xyz=None, embed_dir=None):
        """Build the dexnerf_mlp_program (ctypes struct).  Encoders are only needed by the fused
        query path; forward(x) consumes already-encoded inputs."""
        layers = self._layers()
        if len(layers) > L.MAX_OPS:
            raise ValueError("model has %d layers; the kernel program holds %d" % (len(layers), L.MAX_OPS))
        prog = L.Program()
        prog.n_ops = len(layers)
        prog.dim_xyz, prog.dim_dir = int(self.dim_xyz), int(self.dim_dir)
        dims = {L.ENC_XYZ: self.dim_xyz, L.ENC_DIR: self.dim_dir}
        off, width = 0, 1
        for i, (lin, s0, s1, dst, relu) in enumerate(layers):
            op = prog.ops[i]
            op.src0, op.src0_dim = s0, dims[s0]
            op.src1, op.src1_dim = (s1, dims[s1]) if s1 != L.NONE else (L.NONE, 0)
            if op.src0_dim + op.src1_dim != lin.in_features:
                raise RuntimeError("layer %d: input width %d != in_features %d"
                                   % (i, op.src0_dim + op.src1_dim, lin.in_features))
            op.dst, op.out_dim, op.relu = dst, lin.out_features, int(relu)
            op.w_off = off                                   # every block starts 16-byte aligned
            off = _align4(off + lin.in_features * lin.out_features)
            op.b_off = off
            off = _align4(off + lin.out_features)
            if dst in (L.BUF_A, L.BUF_B):
                dims[dst] = lin.out_features
                width = max(width, lin.out_features)
        prog.max_width = width
        if embed_xyz is not None:
            prog.Lx, prog.include_xyz, prog.log_xyz = (embed_xyz.num_encoding_functions,
                                                       int(embed_xyz.include_input), int(embed_xyz.log_sampling))
        if embed_dir is not None:
            prog.Ld, prog.include_dir, prog.log_dir = (embed_dir.num_encoding_functions,
                                                       int(embed_dir.include_input), int(embed_dir.log_sampling))
        return prog

    def packed_params(self):
        """Flat fp32 device buffer [Wt(in,out) | bias] per layer (every block 16-byte aligned), re-packed
        when a parameter changes.  One kernel launch reads all nn.Linear tensors through a device pointer
        table (rebuilt only when a parameter is re-allocated) and transposes them into a persistent buffer."""
        layers = self._layers()
        key = tuple((lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version)
                    for lin, *_ in layers)
        cache = self.__dict__.get("_packed_cache")
        if cache is None or cache[0] != key:
            with torch.no_grad():
                for lin, *_ in layers:
                    if not lin.weight.is_cuda:
                        raise ValueError("model parameters must live on a CUDA device (call .to('cuda'))")
                    if lin.weight.dtype != torch.float32 or not lin.weight.is_contiguous() or not lin.bias.is_contiguous():
                        raise ValueError("model parameters must be contiguous float32 tensors")
                prog = self.program()
                ptrs = tuple(p for lin, *_ in layers for p in (lin.weight.data_ptr(), lin.bias.data_ptr()))
                table = self.__dict__.get("_ptr_table")
                if table is None or table[0] != ptrs:
                    dev = layers[0][0].weight.device
                    table = (ptrs, torch.tensor(ptrs, dtype=torch.int64).to(dev))
                    self.__dict__["_ptr_table"] = table
                    last = prog.ops[prog.n_ops - 1]
                    self.__dict__["_flat_buf"] = torch.zeros(_align4(last.b_off + last.out_dim), dtype=torch.float32,
                                                             device=dev)
                flat = self.__dict__["_flat_buf"]
                L.check(L.lib().dexnerf_pack_params(prog, L.ptr(table[1]), L.ptr(flat), L.stream_ptr()), "pack_params")
            cache = (key, flat)
            self.__dict__["_packed_cache"] = cache
        return cache[1]

    def invalidate(self):
        """Drop the packed fp32 buffer and the bf16 weight images derived from it.  The caches are keyed on
        (data_ptr, Tensor._version) of every parameter, which covers optimizer steps, `copy_`, `load_state_dict`
        and `.to()`; in-place writes THROUGH `.data` (`p.data.mul_(...)`, EMA or clipping code) do not bump
        `_version` - call this after such an edit.  `load_state_dict` and `_apply` (`.to()`, `.cuda()`, `.half()`...)
        call it themselves."""
        for k in ("_packed_cache", "_tc_cache", "_tc_cache_t"):
            self.__dict__.pop(k, None)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.invalidate()
        return out

    # -- nn.Module API -------------------------------------------------------------------
    def forward(self, x):
        xin = L.dev_f32(x, "x")
        D = self.dim_xyz + self.dim_dir
        if xin.shape[-1] != D:
            raise RuntimeError("expected input with last dim %d (xyz %d + dir %d), got %d"
                               % (D, self.dim_xyz, self.dim_dir, xin.shape[-1]))
        M = xin.numel() // D
        out = torch.empty(xin.shape[:-1] + (4,), dtype=torch.float32, device=xin.device)
        if M:
            prog = self.program()
            L.check(L.lib().dexnerf_mlp_forward(prog, L.ptr(self.packed_params()), L.ptr(xin), M, L.ptr(out),
                                                L.stream_ptr()), type(self).__name__ + ".forward")
        return out


class VeryTinyNeRFModel(_ProgramModule):
    """Three fully connected layers (models.py:4-31)."""

    def __init__(self, filter_size=128, num_encoding_functions=6, use_viewdirs=True):
        super().__init__()
        self.num_encoding_functions = num_encoding_functions
        self.xyz_encoding_dims = 3 + 3 * 2 * num_encoding_functions
        self.viewdir_encoding_dims = 3 + 3 * 2 * num_encoding_functions if use_viewdirs is True else 0
        self.layer1 = _Linear(self.xyz_encoding_dims + self.viewdir_encoding_dims, filter_size)
        self.layer2 = _Linear(filter_size, filter_size)
        self.layer3 = _Linear(filter_size, 4)
        self.relu = torch.nn.functional.relu
        self.dim_xyz, self.dim_dir = self.xyz_encoding_dims, self.viewdir_encoding_dims

    def _layers(self):
        first_src1 = L.ENC_DIR if self.dim_dir else L.NONE
        return [(self.layer1, L.ENC_XYZ, first_src1, L.BUF_A, True),
                (self.layer2, L.BUF_A, L.NONE, L.BUF_B, True),
                (self.layer3, L.BUF_B, L.NONE, L.OUT_ALL, False)]


class MultiHeadNeRFModel(_ProgramModule):
    """Separate density and colour heads (models.py:34-78)."""

    def __init__(self, hidden_size=128, num_encoding_functions=6, use_viewdirs=True):
        super().__init__()
        self.num_encoding_functions = num_encoding_functions
        self.xyz_encoding_dims = 3 + 3 * 2 * num_encoding_functions
        self.viewdir_encoding_dims = 3 + 3 * 2 * num_encoding_functions if use_viewdirs is True else 0
        self.layer1 = _Linear(self.xyz_encoding_dims, hidden_size)
        self.layer2 = _Linear(hidden_size, hidden_size)
        self.layer3_1 = _Linear(hidden_size, 1)
        self.layer3_2 = _Linear(hidden_size, hidden_size)
        self.layer4 = _Linear(self.viewdir_encoding_dims + hidden_size, hidden_size)
        self.layer5 = _Linear(hidden_size, hidden_size)
        self.layer6 = _Linear(hidden_size, 3)
        self.relu = torch.nn.functional.relu
        self.dim_xyz, self.dim_dir = self.xyz_encoding_dims, self.viewdir_encoding_dims

    def _layers(self):
        view = L.ENC_DIR if self.dim_dir else L.NONE
        return [(self.layer1, L.ENC_XYZ, L.NONE, L.BUF_A, True),
                (self.layer2, L.BUF_A, L.NONE, L.BUF_B, True),
                (self.layer3_1, L.BUF_B, L.NONE, L.OUT_SIGMA, False),
                (self.layer3_2, L.BUF_B, L.NONE, L.BUF_A, True),
                (self.layer4, L.BUF_A, view, L.BUF_B, True),
                (self.layer5, L.BUF_B, L.NONE, L.BUF_A, True),
                (self.layer6, L.BUF_A, L.NONE, L.OUT_RGB, False)]


class ReplicateNeRFModel(_ProgramModule):
    """The supplementary-material figure, literally (models.py:81-120)."""

    def __init__(self, hidden_size=256, num_layers=4, num_encoding_fn_xyz=6, num_encoding_fn_dir=4,
                 include_input_xyz=True, include_input_dir=True):
        super().__init__()
        self.dim_xyz = (3 if include_input_xyz else 0) + 2 * 3 * num_encoding_fn_xyz
        self.dim_dir = (3 if include_input_dir else 0) + 2 * 3 * num_encoding_fn_dir
        self.layer1 = _Linear(self.dim_xyz, hidden_size)
        self.layer2 = _Linear(hidden_size, hidden_size)
        self.layer3 = _Linear(hidden_size, hidden_size)
        self.fc_alpha = _Linear(hidden_size, 1)
        self.layer4 = _Linear(hidden_size + self.dim_dir, hidden_size // 2)
        self.layer5 = _Linear(hidden_size // 2, hidden_size // 2)
        self.fc_rgb = _Linear(hidden_size // 2, 3)
        self.relu = torch.nn.functional.relu

    def _layers(self):
        return [(self.layer1, L.ENC_XYZ, L.NONE, L.BUF_A, True),
                (self.layer2, L.BUF_A, L.NONE, L.BUF_B, True),
                (self.fc_alpha, L.BUF_B, L.NONE, L.OUT_SIGMA, False),
                (self.layer3, L.BUF_B, L.NONE, L.BUF_A, False),
                (self.layer4, L.BUF_A, L.ENC_DIR, L.BUF_B, True),
                (self.layer5, L.BUF_B, L.NONE, L.BUF_A, True),
                (self.fc_rgb, L.BUF_A, L.NONE, L.OUT_RGB, False)]


class PaperNeRFModel(_ProgramModule):
    """Fig. 7 of the NeRF paper (models.py:123-182); widths are hard-coded 256/128 and
    num_layers / hidden_size / skip_connect_every are accepted but ignored, as in the reference.
    layers_dir[3] is allocated but unused there too (models.py:158-159 vs :178)."""

    def __init__(self, num_layers=8, hidden_size=256, skip_connect_every=4, num_encoding_fn_xyz=6,
                 num_encoding_fn_dir=4, include_input_xyz=True, include_input_dir=True, use_viewdirs=True):
        super().__init__()
        self.dim_xyz = (3 if include_input_xyz else 0) + 2 * 3 * num_encoding_fn_xyz
        self.dim_dir = (3 if include_input_dir else 0) + 2 * 3 * num_encoding_fn_dir
        self.use_viewdirs = use_viewdirs
        self.layers_xyz = torch.nn.ModuleList()
        self.layers_xyz.append(_Linear(self.dim_xyz, 256))
        for i in range(1, 8):
            self.layers_xyz.append(_Linear(self.dim_xyz + 256 if i == 4 else 256, 256))
        self.fc_feat = _Linear(256, 256)
        self.fc_alpha = _Linear(256, 1)
        self.layers_dir = torch.nn.ModuleList()
        self.layers_dir.append(_Linear(256 + self.dim_dir, 128))
        for i in range(3):
            self.layers_dir.append(_Linear(128, 128))
        self.fc_rgb = _Linear(128, 3)
        self.relu = torch.nn.functional.relu

    def _layers(self):
        if not self.use_viewdirs:
            raise RuntimeError("PaperNeRFModel without view directions has mismatched layer shapes "
                               "in the reference (models.py:157 vs :176)")
        ops, cur, other = [], L.BUF_A, L.BUF_B
        ops.append((self.layers_xyz[0], L.ENC_XYZ, L.NONE, cur, True))
        for i in range(1, 8):
            if i == 4:
                ops.append((self.layers_xyz[i], L.ENC_XYZ, cur, other, True))   # cat((xyz, x))
            else:
                ops.append((self.layers_xyz[i], cur, L.NONE, other, True))
            cur, other = other, cur
        ops.append((self.fc_feat, cur, L.NONE, other, False))
        cur, other = other, cur
        ops.append((self.fc_alpha, cur, L.NONE, L.OUT_SIGMA, False))
        ops.append((self.layers_dir[0], cur, L.ENC_DIR, other, True))
        cur, other = other, cur
        for i in range(1, 3):
            ops.append((self.layers_dir[i], cur, L.NONE, other, True))
            cur, other = other, cur
        ops.append((self.fc_rgb, cur, L.NONE, L.OUT_RGB, False))
        return ops


class FlexibleNeRFModel(_ProgramModule):
    """The model every reference script instantiates (models.py:185-256)."""

    def __init__(self, num_layers=4, hidden_size=128, skip_connect_every=4, num_encoding_fn_xyz=6,
                 num_encoding_fn_dir=4, include_input_xyz=True, include_input_dir=True, use_viewdirs=True):
        super().__init__()
        self.dim_xyz = (3 if include_input_xyz else 0) + 2 * 3 * num_encoding_fn_xyz
        self.dim_dir = (3 if include_input_dir else 0) + 2 * 3 * num_encoding_fn_dir
        self.skip_connect_every = skip_connect_every
        if not use_viewdirs:
            self.dim_dir = 0
        self.layer1 = _Linear(self.dim_xyz, hidden_size)
        self.layers_xyz = torch.nn.ModuleList()
        for i in range(num_layers - 1):
            takes_skip = i % self.skip_connect_every == 0 and i > 0 and i != num_layers - 1
            self.layers_xyz.append(_Linear(self.dim_xyz + hidden_size if takes_skip else hidden_size, hidden_size))
        self.use_viewdirs = use_viewdirs
        if self.use_viewdirs:
            self.layers_dir = torch.nn.ModuleList()
            self.layers_dir.append(_Linear(self.dim_dir + hidden_size, hidden_size // 2))
            self.fc_alpha = _Linear(hidden_size, 1)
            self.fc_rgb = _Linear(hidden_size // 2, 3)
            self.fc_feat = _Linear(hidden_size, hidden_size)
        else:
            self.fc_out = _Linear(hidden_size, 4)
        self.relu = torch.nn.functional.relu
        self.hidden_size = hidden_size

    def _layers(self):
        ops, cur, other = [], L.BUF_A, L.BUF_B
        ops.append((self.layer1, L.ENC_XYZ, L.NONE, cur, False))            # no ReLU (models.py:238)
        for i, lin in enumerate(self.layers_xyz):
            skip = i % self.skip_connect_every == 0 and i > 0
            ops.append((lin, cur, L.ENC_XYZ if skip else L.NONE, other, True))  # cat((x, xyz))
            cur, other = other, cur
        if not self.use_viewdirs:
            ops.append((self.fc_out, cur, L.NONE, L.OUT_ALL, False))
            return ops
        ops.append((self.fc_alpha, cur, L.NONE, L.OUT_SIGMA, False))        # alpha from the trunk
        ops.append((self.fc_feat, cur, L.NONE, other, True))
        ops.append((self.layers_dir[0], other, L.ENC_DIR, cur, True))
        ops.append((self.fc_rgb, cur, L.NONE, L.OUT_RGB, False))
        return ops
