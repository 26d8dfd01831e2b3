"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the header
declares, the Python namespace mirrors the reference's, models lower to the right layer programs
and initialise exactly like the reference's classes, and nothing falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import nerf
from nerf import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dexnerf.h")).read()
    declared = sorted(set(re.findall(r"DEXNERF_API\s+[\w\s\*]+?\b(dexnerf_\w+)\s*\(", hdr)))
    assert len(declared) >= 15
    handle = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), name
    assert sorted(L._SIGS) == declared          # the ctypes table binds exactly the header
    assert L.lib().dexnerf_abi_version() == 2


def test_struct_layouts_match_header():
    assert ctypes.sizeof(L.Op) == 48
    assert ctypes.sizeof(L.Program) == 48 + 48 * L.MAX_OPS
    assert ctypes.sizeof(L.FlexibleSpec) == 48


def test_render_params_layout_matches_a_c_compiler(tmp_path):
    """dexnerf_render_params / dexnerf_model_ref as gcc lays them out from include/dexnerf.h against the ctypes
    mirror in nerf/_lib.py (size and the offset of every field)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    fields = [n for n, _ in L.RenderParams._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "dexnerf.h"', 'int main(void) {',
           'printf("%zu %zu\\n", sizeof(dexnerf_render_params), sizeof(dexnerf_model_ref));']
    src += ['printf("%%zu\\n", offsetof(dexnerf_render_params, %s));' % f for f in fields]
    src += ['return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(L.RenderParams) and int(out[1]) == ctypes.sizeof(L.ModelRef)
    for f, off in zip(fields, out[2:]):
        assert getattr(L.RenderParams, f).offset == int(off), f


def test_namespace_matches_reference():
    # names the reference scripts import (train_dexnerf_rgb.py:15-19, tiny_nerf.py:9)
    for name in ["CfgNode", "get_embedding_function", "get_ray_bundle", "img2mse", "meshgrid_xy", "models",
                 "mse2psnr", "run_one_iter_of_nerf", "cumprod_exclusive", "get_minibatches",
                 "positional_encoding", "sample_pdf", "volume_render_radiance_field", "ndc_rays",
                 "predict_and_render_radiance", "run_network", "FlexibleNeRFModel", "PaperNeRFModel"]:
        assert hasattr(nerf, name), name
    assert nerf.sample_pdf is nerf.sample_pdf_2      # the re-binding must be kept
    for cls in ["VeryTinyNeRFModel", "MultiHeadNeRFModel", "ReplicateNeRFModel", "PaperNeRFModel",
                "FlexibleNeRFModel"]:
        assert isinstance(getattr(nerf.models, cls)(), torch.nn.Module)


def test_no_cpu_fallback():
    with pytest.raises(ValueError):
        nerf.positional_encoding(torch.zeros(4, 3))
    with pytest.raises(ValueError):
        nerf.cumprod_exclusive(torch.ones(2, 3))
    with pytest.raises(ValueError):
        nerf.sample_pdf(torch.zeros(2, 5), torch.zeros(2, 4), 8, det=True)
    with pytest.raises(ValueError):
        nerf.volume_render_radiance_field(torch.zeros(2, 3, 4), torch.zeros(2, 3), torch.zeros(2, 3),
                                          m_thres_cand=[5.0])
    with pytest.raises(ValueError):
        nerf.FlexibleNeRFModel()(torch.zeros(3, 66))
    with pytest.raises(TypeError):     # the reference iterates over None (volume_rendering_utils.py:53)
        nerf.volume_render_radiance_field(torch.zeros(2, 3, 4), torch.zeros(2, 3), torch.zeros(2, 3))


def _digest(module):
    acc, k = 0.0, 1
    for _, p in module.state_dict().items():
        v = p.detach().double().flatten()
        acc += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64)).sum()) * k
        k += 1
    return acc


def test_models_initialise_like_the_reference(golden):
    g = golden("models")
    torch.manual_seed(42)
    m = nerf.FlexibleNeRFModel(num_layers=8, hidden_size=256, skip_connect_every=4,
                               num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    assert list(m.state_dict().keys()) == g["flex8x256_keys"].tolist()
    assert sum(p.numel() for p in m.parameters()) == int(g["flex8x256_nparams"]) == 595844
    assert _digest(m) == float(g["flex8x256_digest"])
    torch.manual_seed(42)
    m = nerf.FlexibleNeRFModel(num_layers=8, hidden_size=128, skip_connect_every=3,
                               num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    assert _digest(m) == float(g["flex8x128s3_digest"])
    torch.manual_seed(42)
    m = nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    assert _digest(m) == float(g["flex4x128_digest"])
    torch.manual_seed(42)
    m = nerf.PaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    assert list(m.state_dict().keys()) == g["paper_keys"].tolist()
    assert _digest(m) == float(g["paper_digest"])
    torch.manual_seed(42)
    m = nerf.FlexibleNeRFModel(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, use_viewdirs=False)
    assert _digest(m) == float(g["flex_noview_digest"])


def test_pretrained_checkpoint_keys_load(golden):
    g = golden("lego_lowres")
    m = nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    sd = {k[len("coarse."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("coarse.")}
    m.load_state_dict(sd)   # strict: names and shapes are the compat contract


def test_program_lowering_flexible_8x256():
    m = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    p = m.program(nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True))
    ops = [p.ops[i] for i in range(p.n_ops)]
    assert p.n_ops == 12 and p.dim_xyz == 63 and p.dim_dir == 27 and p.max_width == 256
    assert (ops[0].src0, ops[0].relu, ops[0].out_dim) == (L.ENC_XYZ, 0, 256)        # layer1, no ReLU
    skips = [i for i, o in enumerate(ops) if o.src1 == L.ENC_XYZ]
    assert skips == [5] and ops[5].src0_dim == 256 and ops[5].src1_dim == 63       # layers_xyz[4] = cat(x, xyz)
    assert ops[8].dst == L.OUT_SIGMA and ops[8].src0 == ops[9].src0                # alpha from the trunk
    assert (ops[10].src1, ops[10].src1_dim, ops[10].out_dim) == (L.ENC_DIR, 27, 128)
    assert ops[11].dst == L.OUT_RGB
    macs = sum((o.src0_dim + o.src1_dim) * o.out_dim for o in ops)
    assert macs == 593408
    assert sum((o.src0_dim + o.src1_dim + 1) * o.out_dim for o in ops) == 595844   # parameters
    # the flat buffer keeps every weight / bias block 16-byte aligned (vector loads and reductions)
    assert all(o.w_off % 4 == 0 and o.b_off % 4 == 0 for o in ops)
    assert ops[-1].b_off + ops[-1].out_dim == 595844 + 3                            # 3 floats of padding (fc_alpha)
    assert (p.Lx, p.Ld, p.include_xyz, p.log_xyz) == (10, 4, 1, 1)


def test_program_lowering_paper():
    m = nerf.PaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    p = m.program()
    ops = [p.ops[i] for i in range(p.n_ops)]
    assert (ops[4].src0, ops[4].src1) == (L.ENC_XYZ, ops[3].dst)                   # cat((xyz, x))
    macs = sum((o.src0_dim + o.src1_dim) * o.out_dim for o in ops)
    assert macs == 626176                                                          # SURVEY.md section 8(a-3')
    for o in ops:
        assert o.dst != o.src0 and o.dst != o.src1


def test_cfgnode_schema():
    cfg = nerf.CfgNode({"experiment": {"id": "x"}, "models": {"coarse": {"type": "FlexibleNeRFModel"}},
                        "nerf": {"use_viewdirs": True, "train": {"num_coarse": 64}, "validation": {"num_coarse": 64}}})
    assert cfg.models.coarse.type == "FlexibleNeRFModel"
    assert not hasattr(cfg.models, "fine") and hasattr(cfg.models, "coarse")
    assert getattr(cfg.nerf, "train").num_coarse == 64
    with pytest.raises(AttributeError):
        cfg.nerf.nope
    import yaml
    assert yaml.safe_load(cfg.dump())["nerf"]["train"]["num_coarse"] == 64
    assert getattr(nerf.models, cfg.models.coarse.type) is nerf.FlexibleNeRFModel


def test_shipped_yaml_configs_load():
    """Every YAML the reference ships (config/*.yml and pretrained/*/config.yml, copied verbatim to
    tests/golden/configs as data fixtures) goes through CfgNode the way the scripts read it
    (train_dexnerf_rgb.py:39-41 and the key list of SURVEY.md section 5): attribute access for every key a script
    consumes, hasattr gates, getattr(cfg.nerf, mode), models constructible by name, dump() round trip."""
    import glob
    import yaml
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "configs", "*.yml")))
    assert len(files) == 14
    for path in files:
        with open(path, "r") as f:
            cfg = nerf.CfgNode(yaml.load(f, Loader=yaml.FullLoader))
        assert isinstance(cfg.experiment.id, str) and cfg.experiment.train_iters > 0, path
        if os.path.basename(path) == "config_default.yml":
            # an older schema (num_encoding_functions, no dataset.type): the scripts fail on it with the
            # AttributeError of cfgnode.py:115-119 - and so does the drop-in
            assert getattr(nerf.models, cfg.models.coarse.type) is nerf.FlexibleNeRFModel
            with pytest.raises(AttributeError):
                cfg.models.coarse.num_encoding_fn_xyz
            continue
        for key in ("validate_every", "save_every", "print_every", "randomseed", "logdir"):
            getattr(cfg.experiment, key)
        if hasattr(cfg.dataset, "type"):          # config/default.yml predates the dataset.type / near / far keys
            assert cfg.dataset.type in ("blender", "llff", "messytable"), path
            assert isinstance(cfg.dataset.no_ndc, bool) and cfg.dataset.far > cfg.dataset.near >= 0, path
        assert hasattr(cfg.models, "fine") and hasattr(cfg.models, "coarse"), path
        for net in (cfg.models.coarse, cfg.models.fine):
            cls = getattr(nerf.models, net.type)                       # train_dexnerf_rgb.py:122
            m = cls(num_encoding_fn_xyz=net.num_encoding_fn_xyz, num_encoding_fn_dir=net.num_encoding_fn_dir,
                    include_input_xyz=net.include_input_xyz, include_input_dir=net.include_input_dir,
                    use_viewdirs=net.use_viewdirs)                     # the five kwargs of :122-128
            assert m.dim_xyz == (3 if net.include_input_xyz else 0) + 6 * net.num_encoding_fn_xyz
        assert cfg.optimizer.type == "Adam" and cfg.optimizer.lr > 0
        assert cfg.scheduler.lr_decay > 0 and 0 < cfg.scheduler.lr_decay_factor <= 1
        assert isinstance(cfg.nerf.use_viewdirs, bool)
        for mode in ("train", "validation"):
            opt = getattr(cfg.nerf, mode)                              # train_utils.py:114
            assert opt.chunksize > 0 and opt.num_coarse > 0 and opt.num_fine > 0, path
            for key in ("perturb", "white_background", "radiance_field_noise_std", "lindisp"):
                getattr(opt, key)
        assert cfg.nerf.train.num_random_rays > 0
        if "messytable" in os.path.basename(path):
            assert cfg.nerf.validation.m_thres > 0          # (two of them say dataset.type: blender)
        with pytest.raises(AttributeError):
            cfg.dataset.no_such_key                                   # cfgnode.py:115-119
        back = yaml.safe_load(cfg.dump())
        assert back["nerf"]["train"]["num_coarse"] == cfg.nerf.train.num_coarse and back["experiment"]["id"] == cfg.experiment.id


def test_linspace_formula_matches_torch():
    """The kernels rebuild torch.linspace in-register (csrc/common.cuh linspace_at); this pins the
    formula they use against torch on the CPU."""
    f32 = np.float32
    for s, e, n in [(0, 1, 64), (0, 1, 128), (0, 1, 256), (0, 1, 24), (2, 6, 64), (0.3, 4, 64), (1, 32, 6), (0, 1, 17)]:
        s32, e32 = f32(s), f32(e)
        step = f32((e32 - s32) / f32(n - 1))
        mine = np.array([f32(float(s32) + float(step) * i) if i < n // 2 else f32(float(e32) - float(step) * (n - 1 - i))
                         for i in range(n)], dtype=f32)
        assert np.array_equal(mine, torch.linspace(s, e, n, dtype=torch.float32).numpy()), (s, e, n)


def test_meshgrid_and_helpers():
    ii, jj = nerf.meshgrid_xy(torch.arange(3), torch.arange(4, 7))
    a, b = np.meshgrid(np.arange(3), np.arange(4, 7), indexing="xy")
    assert np.array_equal(ii.numpy(), a) and np.array_equal(jj.numpy(), b)
    assert [c.shape[0] for c in nerf.get_minibatches(torch.zeros(10, 2), 4)] == [4, 4, 2]
    assert nerf.mse2psnr(0) == 50.0


def test_bench_scene_tables_are_consistent():
    """bench.py's secondary scenes (c3 / c5): sizes as SURVEY.md section 8 states them, and the per-evaluation
    FLOP count follows from the weight shapes of the network the scene builds."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.FLOP_PER_EVAL == 1186816 and bench.FLOP_PER_RAY == 303824896      # C2
    bench.apply_scene("c5")
    assert (bench.H, bench.W, bench.NC, bench.NF) == (720, 1280, 128, 256)
    assert bench.FLOP_PER_EVAL == 1186816 and bench.FLOP_PER_RAY == (128 + 128 + 256) * 1186816
    bench.apply_scene("c3")
    assert (bench.H, bench.W, bench.NC, bench.NF) == (270, 480, 64, 64) and bench.H * bench.W == 129600
    mc, mf = bench.state_dicts()
    assert mc.hidden_size == 128 and len(mc.layers_xyz) == 7 and mc.skip_connect_every == 3
    assert bench.FLOP_PER_EVAL == 2 * sum(p.numel() for k, p in mc.named_parameters() if k.endswith("weight"))
    assert float(mc.fc_alpha.weight.abs().max()) > 10.0          # the x1000 scale that makes sigma cross thresholds


def test_bench_clock_sampler_keeps_the_samples_of_the_timed_window():
    """bench.py's ClockSampler (an `nvidia-smi -lms` process) may be started before the warm-up; stop(window) keeps
    only the samples whose timestamps fall inside the timed region and collects the throttle reasons seen there."""
    import datetime
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    def stamp(t):
        return datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]

    t0 = 1_700_000_000.0
    lines = [
        "%s, 345, 1965, 180.0, 0x0, Not Active, Not Active, Not Active, Not Active" % stamp(t0 - 0.30),   # idle, before
        "%s, 1965, 1965, 700.0, 0x4, Not Active, Not Active, Not Active, Active" % stamp(t0 + 0.01),
        "%s, 1950, 1965, 900.0, 0x4, Not Active, Not Active, Not Active, Active" % stamp(t0 + 0.05),
        "%s, 1935, 1965, 950.0, 0x0, Not Active, Not Active, Not Active, Not Active" % stamp(t0 + 0.09),
        "%s, 400, 1965, 200.0, 0x0, Active, Not Active, Not Active, Not Active" % stamp(t0 + 0.40),        # after
        "garbage line",
    ]

    class FakeProc:
        def terminate(self):
            pass

        def communicate(self, timeout=None):
            return "\n".join(lines) + "\n", ""

    s = bench.ClockSampler(0, period_ms=25)
    s.proc = FakeProc()
    got = s.stop((t0, t0 + 0.10))
    assert got["samples"] == 3 and got["sm_mhz"] == 1950.0 and got["sm_max_mhz"] == 1965.0
    assert got["reasons"] == ["sw_power_cap"]                      # the hw_slowdown sample lies outside the window
    s.proc = FakeProc()
    everything = s.stop()
    assert everything["samples"] == 5 and everything["reasons"] == ["hw_slowdown", "sw_power_cap"]


def test_built_library_carries_the_tensor_core_instructions():
    """The SASS of the built library (cuobjdump, no GPU needed): the MLP kernels issue tcgen05.mma (UTCHMMA) with TMEM
    loads / stores and bulk copies, the weight-gradient GEMM additionally the legacy HMMA + LDSM of its column sums -
    i.e. the product path is the hand-written sm_100a code, not a library or CUDA-core fallback."""
    import collections
    import re
    import shutil
    import subprocess
    from nerf import _lib as L
    if shutil.which("cuobjdump") is None or not os.path.exists(L.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    assert "sm_100a" in out or "sm_100" in out

    def find(fragment):
        hits = [c for name, c in kernels.items() if fragment in name]
        assert hits, fragment
        return hits

    for frag in ("mlp_tc_kernel", "mlp_tc3_kernel", "mlp_tc_bwd_dx_kernel", "mlp_tc_bwd_dw_kernel"):
        for c in find(frag):
            assert c["UTCHMMA"] > 0 and c["UBLKCP"] > 0 and c["SYNCS"] > 0, frag       # tcgen05.mma, cp.async.bulk, mbarriers
    for frag in ("mlp_tc_kernel", "mlp_tc3_kernel", "mlp_tc_bwd_dx_kernel"):
        for c in find(frag):
            assert c["LDTM"] > 0 and c["STTM"] > 0, frag                                # activations stay in TMEM
    for c in find("mlp_tc_bwd_dw_kernel"):
        assert c["HMMA"] > 0 and c["LDSM"] > 0 and c["LDTM"] > 0
