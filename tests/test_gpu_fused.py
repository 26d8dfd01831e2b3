"""The fused render driver (csrc/render.cu: dexnerf_ray_setup / dexnerf_render_fused_fwd / _bwd) on the GPU:
equal to the step-by-step sequence of the individual C-ABI calls, the camera form equal to get_ray_bundle +
run_one_iter_of_nerf, the Philox draws of the setup launch, NDC in nerf.Trainer, and no torch kernel inside a
validation render."""
import ctypes as C

import numpy as np
import pytest
import torch

import nerf
from nerf import _lib as L
from nerf import render
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
THR = [float(m) for m in range(5, 105, 5)]


def make_cfg(nc, nf, near, far, perturb=False, noise_std=0.0, white=False, lindisp=False, no_ndc=True, chunksize=1 << 20):
    mode = dict(chunksize=chunksize, perturb=perturb, num_coarse=nc, num_fine=nf, white_background=white,
                radiance_field_noise_std=noise_std, lindisp=lindisp)
    return nerf.CfgNode(dict(dataset=dict(no_ndc=no_ndc, near=near, far=far),
                             nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))


def pair(layers, hidden, skip, seed=3, boost=300.0):
    torch.manual_seed(seed)
    mc, mf = nerf.FlexibleNeRFModel(layers, hidden, skip, 10, 4), nerf.FlexibleNeRFModel(layers, hidden, skip, 10, 4)
    with torch.no_grad():
        for m in (mc, mf):
            m.fc_alpha.weight.mul_(boost)
            m.fc_alpha.bias.fill_(1.0)
    return mc.cuda(), mf.cuda()


def camera(H, W, f):
    T = O.pose_spherical_world2cam(25.0, -35.0, 4.0)
    K = torch.tensor([[f, 0, W / 2.0], [0, f, H / 2.0], [0, 0, 1]])
    return T.contiguous().cuda(), K.cuda()        # (torch.linalg.inv returns column-major strides)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_fused_equals_stepwise_calls(precision):
    """run_one_iter_of_nerf (one fused call per chunk) against predict_and_render_radiance on the (n, 11) ray matrix
    the reference builds (train_utils.py:222-250) - the same kernels launched one by one from Python.  The only
    arithmetic that differs is the view-direction norm (torch's CUDA norm there, fp64-accumulated in the setup
    kernel), i.e. a last-bit difference of the directions: maps to 2e-5, Dex depths the same sample everywhere."""
    nerf.set_precision(precision)
    try:
        mc, mf = pair(8, 128, 3)
        H, W = 12, 20
        T, K = camera(H, W, 30.0)
        ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
        cfg = make_cfg(64, 64, 2.0, 6.0, white=True)
        with torch.no_grad():
            ro, rd = nerf.get_ray_bundle(H, W, None, T, K)
            fused = nerf.run_one_iter_of_nerf(H, W, 30.0, mc, mf, ro, rd, cfg, mode="validation", encode_position_fn=ex,
                                              encode_direction_fn=ed, m_thres_cand=THR)
            rof, rdf = ro.reshape(-1, 3), rd.reshape(-1, 3)
            vd = rdf / rdf.norm(p=2, dim=-1).unsqueeze(-1)
            rays = torch.cat((rof, rdf, 2.0 * torch.ones_like(rdf[..., :1]), 6.0 * torch.ones_like(rdf[..., :1]), vd), -1)
            step = nerf.predict_and_render_radiance(rays, mc, mf, cfg, mode="validation", encode_position_fn=ex,
                                                    encode_direction_fn=ed, m_thres_cand=THR)
        assert len(fused) == len(step) == 26 and fused[0].shape == (H, W, 3) and fused[6].shape == (H, W)
        for a, b in zip(fused[:6], step[:6]):
            assert float((a.reshape(b.shape) - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))
        da, db = torch.stack(fused[6:]).reshape(20, -1), torch.stack(step[6:])
        assert float(((da - db).abs() <= 1e-6).float().mean()) > 0.999
        assert float((da > 2.0 + 1e-4).float().mean()) > 0.05           # the thresholds are really crossed
    finally:
        nerf.set_precision("bf16")


def test_render_camera_equals_bundle_path_and_chunks():
    """nerf.render_camera generates the rays inside the setup launch: bit-identical to get_ray_bundle +
    run_one_iter_of_nerf, for the whole frame, for a row block, and with a chunk size that splits the frame."""
    mc, mf = pair(8, 256, 4, seed=5, boost=150.0)
    H, W = 16, 24
    T, K = camera(H, W, 30.0)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    kw = dict(mode="validation", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=THR)
    with torch.no_grad():
        ro, rd = nerf.get_ray_bundle(H, W, None, T, K)
        a = nerf.run_one_iter_of_nerf(H, W, 30.0, mc, mf, ro, rd, make_cfg(64, 128, 2.0, 6.0), **kw)
        b = nerf.render_camera(H, W, T, K, mc, mf, make_cfg(64, 128, 2.0, 6.0), **kw)
        c = nerf.render_camera(H, W, T, K, mc, mf, make_cfg(64, 128, 2.0, 6.0, chunksize=100), **kw)   # 4 rows / chunk
        d = nerf.render_camera(H, W, T, K, mc, mf, make_cfg(64, 128, 2.0, 6.0), row_start=5, row_count=7, **kw)
    assert len(a) == len(b) == 26
    for x, y, z_, w in zip(a, b, c, d):
        assert torch.equal(x, y) and torch.equal(x, z_) and torch.equal(x[5:12], w)


def _setup(n, Nc, Nf, perturb, std, offset, seed=1234, replay=None):
    """dexnerf_ray_setup alone; returns views into its workspace."""
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(7)
    ro = torch.randn(n, 3, generator=g).cuda()
    rd = torch.randn(n, 3, generator=g).cuda()
    ws = torch.empty(render.workspace_bytes(n, Nc, Nf), dtype=torch.uint8, device=dev)
    lay = render.workspace_layout(n, Nc, Nf)
    p = L.RenderParams()
    p.n, p.ro, p.rd = n, ro.data_ptr(), rd.data_ptr()
    p.use_viewdirs, p.near, p.far, p.Nc, p.Nf, p.perturb, p.noise_std = 1, 2.0, 6.0, Nc, Nf, int(perturb), std
    p.seed, p.offset = seed, offset
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    if replay is not None:
        p.t_rand = replay.data_ptr()
    L.check(L.lib().dexnerf_ray_setup(C.byref(p), L.stream_ptr()), "ray_setup")
    torch.cuda.synchronize()
    v = lambda name, shape: render.ws_view(ws, lay, name, shape).clone()
    return dict(ro=ro, rd=rd, vd=v("viewdirs", (n, 3)), z=v("z_coarse", (n, Nc)), u=v("u", (n, Nf)),
                noise_c=v("noise_coarse", (n, Nc)), noise_f=v("noise_fine", (n, Nc + Nf)))


def test_ray_setup_depths_viewdirs_and_philox_draws():
    n, Nc, Nf = 2048, 64, 128
    # validation: plain linspace depths (bit-identical to the stand-alone kernel / the oracle), unit view directions
    s = _setup(n, Nc, Nf, False, 0.0, 0)
    lin = torch.linspace(0.0, 1.0, Nc)
    ref = (2.0 * (1.0 - lin) + 6.0 * lin).cuda()              # train_utils.py:118-120 on the CPU, bit for bit
    assert torch.equal(s["z"], ref.expand(n, Nc))
    vd_ref = (s["rd"].double() / s["rd"].double().norm(dim=-1, keepdim=True)).float()
    assert float((s["vd"] - vd_ref).abs().max()) <= 1.2e-7
    # train: jittered depths stay inside their bins and sorted; u uniform on [0, 1); noises ~ N(0, std^2)
    a = _setup(n, Nc, Nf, True, 0.2, 11)
    assert bool((a["z"][:, 1:] >= a["z"][:, :-1]).all()) and float(a["z"].min()) >= 2.0 and float(a["z"].max()) <= 6.0
    assert not torch.equal(a["z"], s["z"])
    u = a["u"]
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 5e-3 and abs(float(u.var()) - 1.0 / 12.0) < 2e-3
    hist = torch.histc(u, bins=16, min=0.0, max=1.0) / u.numel()
    assert float((hist - 1.0 / 16).abs().max()) < 3e-3
    for key in ("noise_c", "noise_f"):
        x = a[key] / 0.2
        assert abs(float(x.mean())) < 1e-2 and abs(float(x.var()) - 1.0) < 2e-2
        assert abs(float((x ** 4).mean()) - 3.0) < 0.15 and float(x.abs().max()) < 7.0    # normal kurtosis, sane tails
    # rows are not correlated with each other or across the three streams
    assert abs(float(torch.corrcoef(torch.stack((u[0::2].reshape(-1), u[1::2].reshape(-1))))[0, 1])) < 1e-2
    assert abs(float(torch.corrcoef(torch.stack((a["noise_c"].reshape(-1), a["noise_f"][:, :Nc].reshape(-1))))[0, 1])) < 1e-2
    # counter-based: same (seed, offset) -> same numbers; another offset or seed -> different numbers
    b = _setup(n, Nc, Nf, True, 0.2, 11)
    c = _setup(n, Nc, Nf, True, 0.2, 12)
    d = _setup(n, Nc, Nf, True, 0.2, 11, seed=99)
    assert torch.equal(a["u"], b["u"]) and torch.equal(a["noise_f"], b["noise_f"]) and torch.equal(a["z"], b["z"])
    assert not torch.equal(a["u"], c["u"]) and not torch.equal(a["u"], d["u"]) and not torch.equal(a["z"], c["z"])
    # a replayed jitter is used as given
    t_rand = torch.rand(n, Nc, generator=torch.Generator().manual_seed(2)).cuda()
    e = _setup(n, Nc, Nf, True, 0.0, 5, replay=t_rand)
    zo = torch.empty(n, Nc, device="cuda")
    L.check(L.lib().dexnerf_stratified_z(n, Nc, 2.0, 6.0, None, None, 0, L.ptr(t_rand), L.ptr(zo), L.stream_ptr()), "z")
    assert torch.equal(e["z"], zo)


def test_validation_render_launches_only_library_kernels():
    """a validation render must not launch torch (at::) kernels - no norm / cat / ones_like /
    contiguous glue.  The profiler lists every kernel of one warm run_one_iter_of_nerf and one render_camera call."""
    from torch.profiler import ProfilerActivity, profile
    mc, mf = pair(8, 256, 4, seed=9)
    H, W = 16, 32
    T, K = camera(H, W, 40.0)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    cfg = make_cfg(64, 128, 2.0, 6.0)
    kw = dict(mode="validation", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=THR)
    with torch.no_grad():
        ro, rd = nerf.get_ray_bundle(H, W, None, T, K)
        nerf.run_one_iter_of_nerf(H, W, 40.0, mc, mf, ro, rd, cfg, **kw)      # warm-up: weights packed, workspace made
        nerf.render_camera(H, W, T, K, mc, mf, cfg, **kw)
        torch.cuda.synchronize()
        n0 = L.launch_count
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            nerf.run_one_iter_of_nerf(H, W, 40.0, mc, mf, ro, rd, cfg, **kw)
            mid = L.launch_count
            nerf.render_camera(H, W, T, K, mc, mf, cfg, **kw)
            torch.cuda.synchronize()
    assert mid - n0 == 6 and L.launch_count - mid == 6
    from torch.autograd import DeviceType
    kernels = [e.key for e in prof.key_averages() if e.device_type == DeviceType.CUDA and "memcpy" not in e.key.lower()
               and "memset" not in e.key.lower()]
    if not kernels:
        pytest.skip("the profiler recorded no device activity in this environment")
    foreign = [k for k in kernels if "dexnerf" not in k]
    assert not foreign, foreign
    total = sum(e.count for e in prof.key_averages() if e.device_type == DeviceType.CUDA and "dexnerf" in e.key)
    assert total == 12, total


def test_trainer_applies_ndc_like_run_one_iter_of_nerf():
    """with cfg.dataset.no_ndc False (the LLFF configs) run_one_iter_of_nerf warps the rays with
    ndc_rays (train_utils.py:238-242); nerf.Trainer must optimise the same parameterisation: same loss and the
    same gradients as the autograd path on the same rays and replayed draws."""
    import copy
    torch.manual_seed(13)
    mk = lambda: nerf.FlexibleNeRFModel(8, 128, 3, 6, 4)
    mc, mf = mk().cuda(), mk().cuda()
    mc2, mf2 = copy.deepcopy(mc), copy.deepcopy(mf)
    ex, ed = nerf.get_embedding_function(6, True, True), nerf.get_embedding_function(4, True, True)
    H, W, focal, n, nc, nf = 12, 16, 20.0, 12 * 16, 16, 24
    g = torch.Generator().manual_seed(6)
    ro = torch.cat((0.1 * torch.randn(n, 2, generator=g), torch.zeros(n, 1)), -1).cuda()      # forward-facing rays
    rd = torch.cat((0.3 * torch.randn(n, 2, generator=g), -torch.ones(n, 1)), -1).cuda()
    target = torch.rand(n, 3, generator=g).cuda()
    rng = dict(t_rand=torch.rand(n, nc, generator=g).cuda(), u=torch.rand(n, nf, generator=g).cuda(),
               noise_coarse=(0.2 * torch.randn(n, nc, generator=g)).cuda(),
               noise_fine=(0.2 * torch.randn(n, nc + nf, generator=g)).cuda())
    cfg = make_cfg(nc, nf, 0.0, 1.0, perturb=True, noise_std=0.2, no_ndc=False)
    out = nerf.run_one_iter_of_nerf(H, W, focal, mc, mf, ro, rd, cfg, mode="train", encode_position_fn=ex,
                                    encode_direction_fn=ed, m_thres_cand=[], rng=rng)
    loss = torch.nn.functional.mse_loss(out[0], target) + torch.nn.functional.mse_loss(out[3], target)
    loss.backward()
    trainer = nerf.Trainer(mc2, mf2, cfg, ex, ed, lr=5e-3)
    trainer.keep_grads = True
    with pytest.raises(nerf.DexNerfError, match="no_ndc"):
        trainer.step(ro, rd, target, rng=rng)
    lt = trainer.step(ro, rd, target, rng=rng, height=H, width=W, focal_length=focal)
    assert abs(float(lt[0]) - float(loss)) < 1e-5
    # ... and NOT the loss of the un-warped rays
    cfg_raw = make_cfg(nc, nf, 0.0, 1.0, perturb=True, noise_std=0.2, no_ndc=True)
    with torch.no_grad():
        raw = nerf.run_one_iter_of_nerf(H, W, focal, mc, mf, ro, rd, cfg_raw, mode="train", encode_position_fn=ex,
                                        encode_direction_fn=ed, m_thres_cand=[], rng=rng)
    raw_loss = torch.nn.functional.mse_loss(raw[0], target) + torch.nn.functional.mse_loss(raw[3], target)
    assert abs(float(raw_loss) - float(loss)) > 1e-4
    from nerf import training
    for i, m in enumerate((mc, mf)):
        flat = trainer._flat(trainer.grads, i)
        got = training.unflatten_grads(m, trainer.progs[i], flat)
        want = [p.grad for lin, *_ in m._layers() for p in (lin.weight, lin.bias)]
        num = sum(float(((a - b) ** 2).sum()) for a, b in zip(got, want))
        den = sum(float((b ** 2).sum()) for b in want)
        assert den > 0 and (num / den) ** 0.5 < 1e-4
    with pytest.raises(nerf.DexNerfError, match="use_viewdirs"):
        bad = make_cfg(nc, nf, 0.0, 1.0)
        bad.nerf.use_viewdirs = False
        nerf.Trainer(mc2, mf2, bad, ex, ed)


def test_fused_backward_single_call_equals_split_calls():
    """dexnerf_render_fused_bwd with which = 3 (both chains in one call) against the Trainer's two calls."""
    torch.manual_seed(2)
    mk = lambda: nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda()
    mc, mf = mk(), mk()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    n, nc, nf = 80, 32, 32
    cfg = make_cfg(nc, nf, 2.0, 6.0, perturb=True, noise_std=0.2)
    g = torch.Generator().manual_seed(4)
    ro = (0.2 * torch.randn(n, 3, generator=g)).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    target = torch.rand(n, 3, generator=g).cuda()
    rng = dict(t_rand=torch.rand(n, nc, generator=g).cuda(), u=torch.rand(n, nf, generator=g).cuda(),
               noise_coarse=(0.2 * torch.randn(n, nc, generator=g)).cuda(),
               noise_fine=(0.2 * torch.randn(n, nc + nf, generator=g)).cuda())
    tr = nerf.Trainer(mc, mf, cfg, ex, ed, lr=0.0)        # lr 0: the Adam step leaves the weights (and their images) alone
    tr.keep_grads = True
    tr.step(ro, rd, target, rng=rng)
    want = tr.grads.clone()
    # replay the backward of the recorded forward in ONE call
    buf = tr._chunk_buffers(n)
    p, keep = tr._render_params(n, buf, ro, rd, rng, None, None, None)
    grads = torch.zeros_like(tr.grads)
    gc, gf = tr._flat(grads, 0), tr._flat(grads, 1)
    L.check(L.lib().dexnerf_render_fused_bwd(C.byref(p), L.ptr(buf["g_rgb"][0]), L.ptr(buf["g_rgb"][1]), L.ptr(buf["d_rf"]),
                                             L.ptr(gc), L.ptr(gf), 3, L.stream_ptr()), "render_fused_bwd")
    torch.cuda.synchronize()
    # same kernels on the same tape; the split-K reduction (red.global.add) orders its fp32 additions differently
    assert float(want.abs().max()) > 0
    assert float((grads - want).norm() / want.norm()) < 1e-5
