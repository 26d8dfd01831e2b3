"""Edge cases of the hot path on the GPU: empty and ragged inputs, single samples, the largest
supported per-ray sizes, non-contiguous / wrong-device arguments."""
import pytest
import torch

import nerf
from nerf import tensorcore
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu


def cfg(nc, nf, chunksize=1 << 20, perturb=False, std=0.0):
    mode = dict(chunksize=chunksize, perturb=perturb, num_coarse=nc, num_fine=nf, white_background=False,
                radiance_field_noise_std=std, lindisp=False)
    return nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=2.0, far=6.0),
                             nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))


def test_empty_inputs():
    z0 = torch.zeros(0, 64, device="cuda")
    res = nerf.volume_render_radiance_field(torch.zeros(0, 64, 4, device="cuda"), z0, torch.zeros(0, 3, device="cuda"),
                                            m_thres_cand=[5.0, 10.0])
    assert len(res) == 7 and res[0].shape == (0, 3) and res[5].shape == (0,)
    assert nerf.sample_pdf(torch.zeros(0, 63, device="cuda"), torch.zeros(0, 62, device="cuda"), 128, det=True).shape == (0, 128)
    assert nerf.cumprod_exclusive(torch.zeros(0, 7, device="cuda")).shape == (0, 7)
    torch.manual_seed(0)
    mc, mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda(), nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    with torch.no_grad():
        out = nerf.run_one_iter_of_nerf(0, 0, 1.0, mc, mf, torch.zeros(0, 3, device="cuda"), torch.zeros(0, 3, device="cuda"),
                                        cfg(64, 128), mode="train", encode_position_fn=ex, encode_direction_fn=ed,
                                        m_thres_cand=[5.0])
    # no rays -> no ray chunks -> the reference's zip(*[]) gives an empty tuple (train_utils.py:252-282); so do we
    assert out == ()


@pytest.mark.parametrize("n,S", [(1, 1), (1, 127), (3, 43), (2, 129), (257, 1)])
def test_ragged_tiles_tensor_core(n, S):
    """Sample counts that do not fill a 128-sample tile / a tile pair: the padded rows must not leak."""
    torch.manual_seed(1)
    model = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    g = torch.Generator().manual_seed(n * 31 + S)
    ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    guard = torch.full((n * S * 4 + 1024,), 7.0, device="cuda")
    rf = guard[: n * S * 4].view(n, S, 4)
    tensorcore.query(model, prog, ro, rd, vd, z, rf)
    assert torch.all(guard[n * S * 4:] == 7.0)                       # nothing written past the last sample
    nerf.set_precision("fp32")
    try:
        ref = nerf.train_utils.query_field(model, ro, rd, vd, z, ex, ed)
    finally:
        nerf.set_precision("bf16")
    assert float((rf - ref).abs().max()) < 3e-2 * max(1.0, float(ref.abs().max()))


def test_largest_per_ray_sizes():
    """S = 2048 samples per ray (the compositing backward's chunk table limit) and 64 thresholds."""
    g = torch.Generator().manual_seed(2)
    n, S, T = 5, 2048, 64
    rf = torch.randn(n, S, 4, generator=g)
    rf[..., 3] = 20 * torch.randn(n, S, generator=g)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values
    rd = torch.randn(n, 3, generator=g)
    thr = [float(x) for x in torch.linspace(1, 64, T)]
    ref = O.volume_render_radiance_field(rf, z, rd, 0.0, False, thr)
    got = nerf.volume_render_radiance_field(rf.cuda(), z.cuda(), rd.cuda(), m_thres_cand=thr)
    for a, b in zip(got[:5], ref[:5]):
        assert torch.allclose(a.cpu(), b, rtol=2e-5, atol=2e-6, equal_nan=True)
    assert torch.equal(torch.stack(got[5:]).cpu(), torch.stack(ref[5:]))
    with pytest.raises(nerf.DexNerfError):
        nerf.volume_render_radiance_field(rf.cuda(), z.cuda(), rd.cuda(), m_thres_cand=list(range(65)))
    from nerf import training
    grad = training.volume_render_backward(rf.cuda(), z.cuda(), rd.cuda(), None, False, torch.ones(n, 3, device="cuda"),
                                           None, None)
    assert torch.isfinite(grad).all()
    with pytest.raises(nerf.DexNerfError):
        training.volume_render_backward(torch.zeros(1, 2049, 4, device="cuda"), torch.zeros(1, 2049, device="cuda"),
                                        torch.ones(1, 3, device="cuda"), None, False, torch.ones(1, 3, device="cuda"),
                                        None, None)


def test_argument_validation():
    with pytest.raises(ValueError):
        nerf.cumprod_exclusive(torch.zeros(2, 3))                    # CPU tensor: there is no CPU path
    with pytest.raises(ValueError):
        nerf.positional_encoding(torch.zeros(4, 3, device="cuda", dtype=torch.float64))
    x = torch.rand(6, 8, device="cuda")
    a = nerf.cumprod_exclusive(x[:, ::2])                            # non-contiguous views are packed, not misread
    assert torch.allclose(a, nerf.cumprod_exclusive(x[:, ::2].contiguous()))
    with pytest.raises(TypeError):
        nerf.volume_render_radiance_field(torch.zeros(1, 2, 4, device="cuda"), torch.zeros(1, 2, device="cuda"),
                                          torch.ones(1, 3, device="cuda"), m_thres_cand=None)


def test_random_shape_sweep_against_oracle():
    """Seeded sweep over ray / sample / threshold counts that straddle every kernel variant (1, 2 or 4 samples per
    lane in the compositing kernel, partial chunks, merge-by-rank vs sorting network): forward maps, Dex depths,
    resampled depths and the compositing backward against the oracle."""
    import random
    from nerf import _lib as L, training
    rnd = random.Random(2026)
    for case in range(24):
        n = rnd.choice([1, 2, 31, 33, 64, 257])
        S = rnd.choice([1, 2, 7, 31, 32, 33, 64, 65, 100, 192, 255, 256, 257, 384, 500])
        T = rnd.choice([0, 1, 3, 20, 33, 64])
        white = rnd.random() < 0.5
        g = torch.Generator().manual_seed(case)
        rf = torch.randn(n, S, 4, generator=g)
        rf[..., 3] = 25 * torch.randn(n, S, generator=g)
        z = torch.sort(0.5 + 5 * torch.rand(n, S, generator=g), dim=-1).values
        rd = torch.randn(n, 3, generator=g)
        noise = 0.3 * torch.randn(n, S, generator=g) if rnd.random() < 0.5 else None
        thr = sorted(rnd.sample(range(1, 120), T))
        rnd.shuffle(thr)                                            # unsorted candidates are ranked in the kernel
        thr = [float(x) for x in thr]
        ref = O.volume_render_radiance_field(rf, z, rd, 0.0, white, thr, noise=noise)
        got = nerf.volume_render_radiance_field(rf.cuda(), z.cuda(), rd.cuda(), 0.0, white, thr,
                                                noise=None if noise is None else noise.cuda())
        tag = (case, n, S, T, white)
        for a, b in zip(got[:5], ref[:5]):
            assert torch.allclose(a.cpu(), b, rtol=3e-5, atol=3e-6, equal_nan=True), tag
        if T:
            assert torch.equal(torch.stack(got[5:]).cpu(), torch.stack(ref[5:])), tag
        # backward
        grgb, gd, ga = torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
        rfo = rf.clone().requires_grad_(True)
        r2 = O.volume_render_radiance_field(rfo, z, rd, 0.0, white, [], noise=noise)
        ((r2[0] * grgb).sum() + (r2[4] * gd).sum() + (r2[2] * ga).sum()).backward()
        bw = training.volume_render_backward(rf.cuda(), z.cuda(), rd.cuda(), None if noise is None else noise.cuda(), white,
                                             grgb.cuda(), gd.cuda(), ga.cuda()).cpu()
        scale = float(rfo.grad.abs().max()) + 1e-12
        assert float((bw - rfo.grad).abs().max()) <= 5e-5 * scale, tag
        # resampling (needs at least 3 coarse depths)
        if S >= 3:
            Nf = rnd.choice([1, 5, 32, 64, 100, 128, 256])
            w = torch.rand(n, S, generator=g) ** 6
            for u in (None, torch.rand(n, Nf, generator=g)):
                mids = 0.5 * (z[:, 1:] + z[:, :-1])
                want = O.merge_fine(z, O.sample_pdf(mids, w[:, 1:-1], Nf, det=(u is None), u=u))
                out = torch.empty(n, S + Nf, device="cuda")
                zc, wc, uc = z.cuda(), w.cuda(), None if u is None else u.cuda()
                L.check(L.lib().dexnerf_resample_merge(L.ptr(zc), L.ptr(wc), n, S, Nf, L.ptr(uc), L.ptr(out), L.stream_ptr()), "rm")
                assert torch.equal(out.cpu(), want), tag + (Nf, u is None)
