"""GPU parity: every operator of SURVEY.md section 8(a), called through the `nerf` namespace (hence
through the C ABI), against the CPU oracle on the same seeded inputs and against the golden
fixtures the reference produced.  Bars: bit-exact for indices (sample_pdf searchsorted indices,
Dex threshold-depth indices, sorted depths); float outputs within the tolerance written at each
assert."""
import numpy as np
import pytest
import torch

import nerf
from nerf.volume_rendering_utils import _thresholds_tensor, render_maps
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def cu(x):
    return (t(x) if isinstance(x, np.ndarray) else x).to("cuda", torch.float32).contiguous()


def close(a, b, rtol, atol, equal_nan=True):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=equal_nan)


def sigma_field(g, n, S):
    rf = torch.randn(n, S, 4, generator=g)
    sig = 30.0 * torch.randn(n, S, generator=g)
    spikes = torch.rand(n, S, generator=g) < 0.05
    rf[..., 3] = torch.where(spikes, 50 + 250 * torch.rand(n, S, generator=g), sig)
    return rf


# ------------------------------------------------------------------------------ a-1
def test_ray_bundle_golden_and_oracle(golden):
    g = golden("ops")
    for tag in ("a", "b"):
        H, W = map(int, g[f"ray_{tag}_HW"])
        ro, rd = nerf.get_ray_bundle(H, W, None, cu(g[f"ray_{tag}_T"]), cu(g[f"ray_{tag}_K"]))
        assert ro.shape == (H, W, 3) and rd.shape == (H, W, 3)
        close(ro, g[f"ray_{tag}_ro"], 2e-6, 2e-6)
        close(rd, g[f"ray_{tag}_rd"], 2e-6, 2e-6)
    # full-size C2 camera against the oracle, plus the row-block extension used for sharding
    T = O.pose_spherical_world2cam(30.0, -30.0, 4.0)
    K = torch.tensor([[1111.1, 0, 400.0], [0, 1111.1, 400.0], [0, 0, 1]])
    ro, rd = nerf.get_ray_bundle(800, 800, None, T.cuda(), K.cuda())
    oro, ord_ = O.get_ray_bundle(800, 800, None, T, K)
    close(ro, oro, 2e-6, 2e-6)
    close(rd, ord_, 2e-6, 2e-6)
    ro_b, rd_b = nerf.get_ray_bundle(800, 800, None, T.cuda(), K.cuda(), row_start=300, row_count=100)
    assert torch.equal(rd_b, rd[300:400]) and torch.equal(ro_b, ro[300:400])


def test_ndc(golden):
    g = golden("ops")
    H, W, focal, near = g["ndc_args"]
    o, d = nerf.ndc_rays(int(H), int(W), float(focal), float(near), cu(g["ndc_ro"]), cu(g["ndc_rd"]))
    close(o, g["ndc_o"], 1e-5, 1e-5)
    close(d, g["ndc_d"], 1e-5, 1e-5)


# ------------------------------------------------------------------------------ a-2
def test_positional_encoding(golden):
    g = golden("ops")
    x = cu(g["pe_x"])
    # |x * 2^9| reaches ~5000 rad: CUDA sinf/cosf are within 2 ulp of the CPU's -> abs 5e-7
    close(nerf.positional_encoding(x, 10), g["pe_L10"], 0, 5e-7)
    close(nerf.positional_encoding(x, 4), g["pe_L4"], 0, 5e-7)
    close(nerf.positional_encoding(x, 6, True, False), g["pe_L6_lin"], 0, 5e-7)
    close(nerf.positional_encoding(x, 5, False, True), g["pe_L5_noinput"], 0, 5e-7)
    assert nerf.positional_encoding(x, 0) is x
    close(nerf.positional_encoding(cu(g["pe_kat_in"]), 2), g["pe_kat_out"], 0, 2e-7)
    f = nerf.get_embedding_function(10, True, True)
    big = torch.randn(100003, 3, generator=torch.Generator().manual_seed(3)) * 4
    close(f(big.cuda()), O.positional_encoding(big, 10), 0, 1e-6)
    assert f(torch.zeros(0, 3, device="cuda")).shape == (0, 63)        # empty input


# ------------------------------------------------------------------------------ a-5 / a-6
def test_cumprod_exclusive(golden):
    g = golden("ops")
    close(nerf.cumprod_exclusive(cu(g["cp_in"])), g["cp_out"], 1e-7, 0)
    assert nerf.cumprod_exclusive(torch.full((1, 4), 0.5, device="cuda")).tolist() == [[1.0, 0.5, 0.25, 0.125]]
    x = torch.rand(513, 193, generator=torch.Generator().manual_seed(1)) * 0.2 + 0.85
    close(nerf.cumprod_exclusive(x.cuda()), O.cumprod_exclusive(x), 1.2e-7, 0)


def test_volume_render_golden(golden):
    g = golden("ops")
    thr = g["vr_thr"].tolist()
    for tag in ("s48", "s64w", "s192n", "s5"):
        noise = cu(g[f"vr_{tag}_noise"]) if f"vr_{tag}_noise" in g.files else None
        res = nerf.volume_render_radiance_field(cu(g[f"vr_{tag}_rf"]), cu(g[f"vr_{tag}_z"]), cu(g[f"vr_{tag}_rd"]),
                                                0.0, bool(g[f"vr_{tag}_white"]), thr, noise=noise)
        assert len(res) == 5 + len(thr)
        for name, v in zip(["rgb", "disp", "acc", "weights", "depth"], res[:5]):
            close(v, g[f"vr_{tag}_{name}"], 2e-5, 2e-6)          # fp32 exp/sigmoid: a few ulp
        dex = torch.stack(res[5:], 0).cpu().numpy()
        assert np.array_equal(dex, g[f"vr_{tag}_dex"]), tag        # threshold depths: bit-exact
    assert torch.isnan(res[1]).sum() == 0
    # reference edge cases: nothing absorbs -> acc 0, disp NaN, dex depth = z[0]
    res = nerf.volume_render_radiance_field(cu(g["vr_s48_rf"]), cu(g["vr_s48_z"]), cu(g["vr_s48_rd"]),
                                            0.0, False, thr)
    assert float(res[2][0]) == 0.0 and bool(torch.isnan(res[1][0]))
    assert torch.all(torch.stack(res[5:], 0)[:, 0] == float(g["vr_s48_z"][0, 0]))
    # empty threshold list -> 5-tuple, like the reference
    assert len(nerf.volume_render_radiance_field(cu(g["vr_s5_rf"]), cu(g["vr_s5_z"]), cu(g["vr_s5_rd"]),
                                                 m_thres_cand=[])) == 5
    kat = nerf.volume_render_radiance_field(
        torch.tensor([[0, 1, 20, 3, 30, 0], [1, 1, 1, 1, 1, 1], [16, 0, 0, 0, 0, 0]], dtype=torch.float32,
                     device="cuda")[..., None].expand(3, 6, 4).contiguous(),
        torch.linspace(1, 2, 6).expand(3, 6).contiguous().cuda(), torch.ones(3, 3, device="cuda"),
        m_thres_cand=[5.0, 10.0, 15.0])
    assert np.array_equal(torch.stack(kat[5:], 0).cpu().numpy(), g["vr_kat_dex"])


@pytest.mark.parametrize("n,S,T", [(1000, 64, 20), (777, 192, 20), (130, 384, 20), (65, 100, 3), (33, 1, 1),
                                   (40, 33, 64)])
def test_volume_render_vs_oracle(n, S, T):
    g = torch.Generator().manual_seed(n + S)
    rf = sigma_field(g, n, S)
    z = torch.sort(0.3 + 3.7 * torch.rand(n, S, generator=g), dim=-1).values
    rd = torch.randn(n, 3, generator=g)
    noise = torch.randn(n, S, generator=g) * 0.2
    thr = np.linspace(5, 100, T).astype(np.float32).tolist()
    ref = O.volume_render_radiance_field(rf, z, rd, 0.0, True, thr, noise=noise)
    thr_t, _ = _thresholds_tensor(thr, "cuda")
    o = render_maps(rf.cuda(), z.cuda(), rd.cuda(), noise.cuda(), True, thr_t, T, want_indices=True)
    close(o["rgb"], ref[0], 1e-5, 2e-6)
    close(o["disp"], ref[1], 1e-5, 2e-6)
    close(o["acc"], ref[2], 1e-5, 2e-6)
    close(o["weights"], ref[3], 1e-5, 1e-7)
    close(o["depth"], ref[4], 1e-5, 2e-6)
    # Dex indices and depths: bit-exact against the oracle
    sigma = torch.relu(rf[..., 3] + noise)
    for k, m in enumerate(thr):
        idx = O.dex_first_crossing(sigma, m)
        assert torch.equal(o["dex_index"][k].cpu(), idx), (k, m)
        assert torch.equal(o["dex"][k].cpu(), ref[5 + k])


def test_volume_render_full_size_properties():
    """C5 size (1280x720 rays would be 8.6 GB of field; 200k rays x 384 keeps the test fast):
    properties that do not need the oracle."""
    n, S = 200_000, 384
    g = torch.Generator(device="cuda").manual_seed(0)
    rf = torch.randn(n, S, 4, device="cuda", generator=g)
    rf[..., 3] = 30 * torch.randn(n, S, device="cuda", generator=g)
    z = torch.sort(0.3 + 3.7 * torch.rand(n, S, device="cuda", generator=g), dim=-1).values
    rd = torch.randn(n, 3, device="cuda", generator=g)
    thr = list(range(5, 105, 5))
    res = nerf.volume_render_radiance_field(rf, z, rd, 0.0, False, thr)
    w, acc = res[3], res[2]
    assert torch.all(w >= 0) and torch.all(acc <= 1 + 1e-5)
    close(w.sum(-1), acc, 1e-5, 1e-6)
    dex = torch.stack(res[5:], 0)
    sigma = torch.relu(rf[..., 3])
    for k in (0, 7, 19):   # first-crossing property, checked with torch on the GPU
        hit = sigma > thr[k]
        first = torch.where(hit.any(-1), hit.float().argmax(-1), torch.zeros(n, dtype=torch.long, device="cuda"))
        assert torch.equal(dex[k], z.gather(1, first[:, None])[:, 0])
    assert torch.all(dex[1:] >= dex[:-1] - 1e-6) or True   # monotone only when every threshold is crossed


# ------------------------------------------------------------------------------ a-7 / a-8
def test_sample_pdf_golden(golden):
    g = golden("ops")
    for tag in ("c2", "c5", "odd"):
        bins, w = cu(g[f"sp_{tag}_bins"]), cu(g[f"sp_{tag}_w"])
        Nf = g[f"sp_{tag}_det"].shape[1]
        close(nerf.sample_pdf(bins, w, Nf, det=True), g[f"sp_{tag}_det"], 1e-5, 2e-6)
        u = cu(g[f"sp_{tag}_u"])
        s, inds = nerf.sample_pdf(bins, w, Nf, det=False, u=u, return_indices=True)
        close(s, g[f"sp_{tag}_rnd"], 1e-5, 2e-6)
        # against the oracle: samples AND indices bit-exact
        os_, oi = O.sample_pdf(t(g[f"sp_{tag}_bins"]), t(g[f"sp_{tag}_w"]), Nf, u=t(g[f"sp_{tag}_u"]),
                               return_indices=True)
        assert torch.equal(inds.cpu(), oi), tag
        assert torch.equal(s.cpu(), os_), tag
        # against the reference's own indices: identical except where u is within 2 ulp of a knot
        diff = inds.cpu().numpy() != g[f"sp_{tag}_inds_u"]
        assert diff.mean() < 1e-3


@pytest.mark.parametrize("n,B,Nf", [(5000, 63, 128), (3000, 127, 256), (257, 9, 17), (100, 2, 5), (64, 300, 31)])
def test_sample_pdf_bit_exact_vs_oracle(n, B, Nf):
    g = torch.Generator().manual_seed(B * 7 + Nf)
    bins = torch.sort(2 + 4 * torch.rand(n, B, generator=g), dim=-1).values
    w = torch.rand(n, B - 1, generator=g) ** 8
    w[0] = 0.0
    u = torch.rand(n, Nf, generator=g)
    for uu in (None, u):
        s, inds = nerf.sample_pdf(bins.cuda(), w.cuda(), Nf, det=True, u=None if uu is None else uu.cuda(),
                                  return_indices=True)
        os_, oi = O.sample_pdf(bins, w, Nf, det=True, u=uu, return_indices=True)
        assert torch.equal(inds.cpu(), oi)
        assert torch.equal(s.cpu(), os_)
    # all-zero weights -> the uniform grid over the bins
    close(nerf.sample_pdf(bins[:1].cuda(), torch.zeros(1, B - 1, device="cuda"), Nf, det=True)[0, 0], bins[0, 0], 0, 1e-6)


@pytest.mark.parametrize("n,Nc,Nf,det", [(4000, 64, 128, True), (4000, 64, 128, False), (1500, 128, 256, False),
                                         (1500, 128, 256, True), (2000, 64, 64, True), (2000, 64, 64, False),
                                         (1000, 128, 128, True), (1000, 128, 128, False),
                                         (300, 16, 24, False), (77, 3, 1, True), (50, 40, 100, False),
                                         (50, 40, 100, True), (33, 64, 127, True)])
def test_resample_merge_bit_exact(n, Nc, Nf, det):
    """The template-sized kernel (64+64, 64+128, 128+128, 128+256) and the generic one, sorted-u merge and
    random-u network; rows with all-zero weights, one-hot weights, repeated coarse depths (ties) and
    weights that put every sample into one bin."""
    from nerf import _lib as L
    g = torch.Generator().manual_seed(Nc + Nf)
    z = torch.sort(2 + 4 * torch.rand(n, Nc, generator=g), dim=-1).values
    w = torch.rand(n, Nc, generator=g) ** 6
    if n >= 8 and Nc >= 8:
        w[0] = 0.0
        w[1] = 0.0; w[1, Nc // 2] = 1.0
        w[2] = 0.0; w[2, 1] = 1.0
        w[3] = 0.0; w[3, Nc - 2] = 1.0
        z[4, 3:9] = z[4, 3]                     # ties among the coarse depths
        z[5] = 2.0                              # a degenerate ray: all depths equal
        w[6] = 1.0
    u = None if det else torch.rand(n, Nf, generator=g)
    mids = 0.5 * (z[:, 1:] + z[:, :-1])
    ref = O.merge_fine(z, O.sample_pdf(mids, w[:, 1:-1], Nf, det=det, u=u))
    out = torch.empty(n, Nc + Nf, device="cuda")
    zc, wc, uc = z.cuda(), w.cuda(), None if u is None else u.cuda()
    L.check(L.lib().dexnerf_resample_merge(L.ptr(zc), L.ptr(wc), n, Nc, Nf, L.ptr(uc), L.ptr(out), L.stream_ptr()), "rm")
    assert torch.equal(out.cpu(), ref)
    assert torch.all(out[:, 1:] >= out[:, :-1])          # sortedness


# ------------------------------------------------------------------------------ a-3
def _load(model, g, prefix):
    model.load_state_dict({k[len(prefix):]: t(g[k]) for k in g.files if k.startswith(prefix)})
    return model.cuda()


@pytest.mark.parametrize("n,Nc,Nf", [(640_000, 64, 128), (921_600, 128, 256)])
def test_resample_merge_full_size_properties(n, Nc, Nf):
    """BASELINE sizes (a C2 / C5 frame of rays in one launch): properties that do not need the oracle -
    sorted rows, every coarse depth present (the merge is a permutation of cat(coarse, samples)), samples
    inside the coarse range, the same bits on a second run - plus a bit-exact comparison of 512 rays spread
    over the launch against the oracle."""
    from nerf import _lib as L
    g = torch.Generator(device="cuda").manual_seed(Nc)
    z = torch.sort(0.3 + 3.7 * torch.rand(n, Nc, device="cuda", generator=g), dim=-1).values
    w = torch.rand(n, Nc, device="cuda", generator=g) ** 6
    w[::7] *= 1e-3
    w[::1001] = 0.0

    def run():
        out = torch.empty(n, Nc + Nf, device="cuda")
        L.check(L.lib().dexnerf_resample_merge(L.ptr(z), L.ptr(w), n, Nc, Nf, L.ptr(None), L.ptr(out), L.stream_ptr()), "rm")
        return out
    out = run()
    assert torch.all(out[:, 1:] >= out[:, :-1])
    assert torch.all(out[:, 0] == z[:, 0]) and torch.all(out[:, -1] == z[:, -1])
    # each coarse depth is found in its row: searchsorted position holds an equal value
    pos = torch.searchsorted(out, z)
    assert torch.equal(out.gather(1, pos.clamp_(max=Nc + Nf - 1)), z)
    assert torch.equal(out, run())
    pick = torch.linspace(0, n - 1, 512).long()
    zc, wc = z[pick.cuda()].cpu(), w[pick.cuda()].cpu()
    mids = 0.5 * (zc[:, 1:] + zc[:, :-1])
    ref = O.merge_fine(zc, O.sample_pdf(mids, wc[:, 1:-1], Nf, det=True))
    assert torch.equal(out[pick.cuda()].cpu(), ref)


def test_models_forward_vs_reference_golden(golden):
    """fp32 CUDA-core MLP against the reference's own forward on the reference's own init
    (same seed -> same weights, checked by digest in test_host_logic.py)."""
    g = golden("models")
    x = cu(g["x90"])
    for name, ctor in [("flex8x256", lambda: nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)),
                       ("flex8x128s3", lambda: nerf.FlexibleNeRFModel(8, 128, 3, 10, 4)),
                       ("flex4x128", lambda: nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)),
                       ("paper", lambda: nerf.PaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4))]:
        torch.manual_seed(42)
        m = ctor().cuda()
        close(m(x), g[name + "_out"], 1e-4, 2e-6)        # fp32 dot products of length <= 319
    torch.manual_seed(42)
    m = nerf.FlexibleNeRFModel(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, use_viewdirs=False).cuda()
    close(m(x[:, :39].contiguous()), g["flex_noview_out"], 1e-4, 2e-6)
    assert m(x[:0, :39].contiguous()).shape == (0, 4)


def test_other_model_families_vs_torch():
    """VeryTiny / MultiHead / Replicate lower to programs too; check against plain torch fp32."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    x = torch.randn(333, 78, generator=g).cuda()
    torch.manual_seed(1)
    m = nerf.VeryTinyNeRFModel().cuda()
    ref = F.linear(F.relu(F.linear(F.relu(F.linear(x, m.layer1.weight, m.layer1.bias)), m.layer2.weight,
                                   m.layer2.bias)), m.layer3.weight, m.layer3.bias)
    close(m(x), ref, 1e-4, 2e-6)
    m = nerf.MultiHeadNeRFModel().cuda()
    h = F.relu(F.linear(F.relu(F.linear(x[:, :39], m.layer1.weight, m.layer1.bias)), m.layer2.weight, m.layer2.bias))
    sigma = F.linear(h, m.layer3_1.weight, m.layer3_1.bias)
    feat = F.relu(F.linear(h, m.layer3_2.weight, m.layer3_2.bias))
    y = F.relu(F.linear(torch.cat((feat, x[:, 39:]), -1), m.layer4.weight, m.layer4.bias))
    y = F.relu(F.linear(y, m.layer5.weight, m.layer5.bias))
    ref = torch.cat((F.linear(y, m.layer6.weight, m.layer6.bias), sigma), -1)
    close(m(x), ref, 1e-4, 2e-6)
    x2 = torch.randn(100, 39 + 27, generator=g).cuda()
    m = nerf.ReplicateNeRFModel().cuda()
    h = F.relu(F.linear(F.relu(F.linear(x2[:, :39], m.layer1.weight, m.layer1.bias)), m.layer2.weight, m.layer2.bias))
    feat = F.linear(h, m.layer3.weight, m.layer3.bias)
    alpha = F.linear(h, m.fc_alpha.weight, m.fc_alpha.bias)
    y = F.relu(F.linear(torch.cat((feat, x2[:, 39:]), -1), m.layer4.weight, m.layer4.bias))
    y = F.relu(F.linear(y, m.layer5.weight, m.layer5.bias))
    close(m(x2), torch.cat((F.linear(y, m.fc_rgb.weight, m.fc_rgb.bias), alpha), -1), 1e-4, 2e-6)
