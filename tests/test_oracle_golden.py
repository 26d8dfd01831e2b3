"""The CPU oracle (oracle/nerf_oracle.py) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import nerf_oracle as O

t = torch.from_numpy


def close(a, b, rtol=2e-6, atol=2e-6, equal_nan=True):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=equal_nan)


def test_ray_bundle(golden):
    g = golden("ops")
    for tag in ("a", "b"):
        H, W = g[f"ray_{tag}_HW"]
        ro, rd = O.get_ray_bundle(int(H), int(W), None, t(g[f"ray_{tag}_T"]), t(g[f"ray_{tag}_K"]))
        assert ro.shape == (H, W, 3)
        close(ro, g[f"ray_{tag}_ro"], 1e-6, 1e-6)
        close(rd, g[f"ray_{tag}_rd"], 1e-6, 1e-6)


def test_ndc(golden):
    g = golden("ops")
    H, W, focal, near = g["ndc_args"]
    o, d = O.ndc_rays(int(H), int(W), float(focal), float(near), t(g["ndc_ro"]), t(g["ndc_rd"]))
    close(o, g["ndc_o"], 1e-5, 1e-5)
    close(d, g["ndc_d"], 1e-5, 1e-5)


def test_positional_encoding(golden):
    g = golden("ops")
    x = t(g["pe_x"])
    assert np.array_equal(O.positional_encoding(x, 10).numpy(), g["pe_L10"])
    assert np.array_equal(O.positional_encoding(x, 4).numpy(), g["pe_L4"])
    assert np.array_equal(O.positional_encoding(x, 6, True, False).numpy(), g["pe_L6_lin"])
    assert np.array_equal(O.positional_encoding(x, 5, False, True).numpy(), g["pe_L5_noinput"])
    assert np.array_equal(O.positional_encoding(x, 0).numpy(), g["pe_L0"])
    k = O.positional_encoding(t(g["pe_kat_in"]), 2).numpy()
    assert np.array_equal(k, g["pe_kat_out"])
    s, c = np.sin, np.cos
    close(k[0], [1, 2, 3, s(1), s(2), s(3), c(1), c(2), c(3), s(2), s(4), s(6), c(2), c(4), c(6)], 1e-6, 1e-6)


def test_cumprod_exclusive(golden):
    g = golden("ops")
    close(O.cumprod_exclusive(t(g["cp_in"])), g["cp_out"], 1e-7, 0)
    assert np.array_equal(O.cumprod_exclusive(torch.tensor([[.5, .5, .5, .5]])).numpy(), g["cp_kat"])
    assert g["cp_kat"].tolist() == [[1.0, 0.5, 0.25, 0.125]]


def test_volume_render(golden):
    g = golden("ops")
    thr = g["vr_thr"].tolist()
    for tag in ("s48", "s64w", "s192n", "s5"):
        noise = t(g[f"vr_{tag}_noise"]) if f"vr_{tag}_noise" in g.files else None
        res = O.volume_render_radiance_field(t(g[f"vr_{tag}_rf"]), t(g[f"vr_{tag}_z"]), t(g[f"vr_{tag}_rd"]),
                                             0.0, bool(g[f"vr_{tag}_white"]), thr, noise=noise)
        for name, v in zip(["rgb", "disp", "acc", "weights", "depth"], res[:5]):
            close(v, g[f"vr_{tag}_{name}"], 2e-5, 2e-6)
        dex = np.stack([v.numpy() for v in res[5:]], 0)
        assert np.array_equal(dex, g[f"vr_{tag}_dex"]), tag   # threshold depths: bit-exact
    # edge rows of s48: nothing absorbs -> acc 0, disp NaN, dex = z[0]
    assert g["vr_s48_acc"][0] == 0 and np.isnan(g["vr_s48_disp"][0])
    assert np.all(g["vr_s48_dex"][:, 0] == g["vr_s48_z"][0, 0])
    assert abs(g["vr_s48_acc"][2] - 1.0) < 1e-6       # last sample absorbs everything


def test_dex_known_answers(golden):
    g = golden("ops")
    sig = torch.tensor([[0, 1, 20, 3, 30, 0], [1, 1, 1, 1, 1, 1], [16, 0, 0, 0, 0, 0]], dtype=torch.float32)
    rf = torch.zeros(3, 6, 4)
    rf[..., 3] = sig
    z = torch.linspace(1, 2, 6).expand(3, 6).contiguous()
    res = O.volume_render_radiance_field(rf, z, torch.ones(3, 3), 0.0, False, [5.0, 10.0, 15.0])
    dex = np.stack([v.numpy() for v in res[5:]], 0)
    assert np.array_equal(dex, g["vr_kat_dex"])
    close(dex[:, 0], [1.4, 1.4, 1.4], 1e-6)
    close(dex[:, 1], [1.0, 1.0, 1.0], 0, 0)
    close(dex[:, 2], [1.0, 1.0, 1.0], 0, 0)
    assert O.dex_first_crossing(sig, 15.0).tolist() == [2, 0, 0]
    assert O.dex_first_crossing(sig, 25.0).tolist() == [4, 0, 0]


def test_sample_pdf(golden):
    g = golden("ops")
    for tag in ("c2", "c5", "odd"):
        bins, w = t(g[f"sp_{tag}_bins"]), t(g[f"sp_{tag}_w"])
        Nf = g[f"sp_{tag}_det"].shape[1]
        assert np.array_equal(g[f"sp_{tag}_det"], g[f"sp_{tag}_det_v1"])      # v1 == v2 in the reference
        det = O.sample_pdf(bins, w, Nf, det=True)
        close(det, g[f"sp_{tag}_det"], 1e-5, 2e-6)
        u = t(g[f"sp_{tag}_u"])
        rnd, inds = O.sample_pdf(bins, w, Nf, det=False, u=u, return_indices=True)
        close(rnd, g[f"sp_{tag}_rnd"], 1e-5, 2e-6)
        # the (cdf, u) -> inds boundary: exact on the reference's own cdf ...
        ref_cdf = t(g[f"sp_{tag}_cdf"])
        assert np.array_equal(O.searchsorted_right(ref_cdf, u).numpy(), g[f"sp_{tag}_inds_u"])
        # ... and our fp64-accumulated cdf is within 2 ulp of the reference's, so indices can only
        # differ where u sits within that distance of a cdf knot.
        cdf = O.pdf_to_cdf(w)
        close(cdf, ref_cdf, 0, 2.5e-7)
        diff = inds.numpy() != g[f"sp_{tag}_inds_u"]
        if diff.any():
            r, c = np.nonzero(diff)
            for rr, cc in zip(r, c):
                k = min(int(inds[rr, cc]), int(g[f"sp_{tag}_inds_u"][rr, cc]))
                assert abs(float(u[rr, cc]) - float(ref_cdf[rr, k])) <= 2.5e-7
        assert diff.mean() < 1e-3
    # all-zero weights -> uniform over the bins
    bins = torch.linspace(2, 6, 9)[None]
    out = O.sample_pdf(bins, torch.zeros(1, 8), 17, det=True)
    close(out, torch.linspace(2, 6, 17)[None], 1e-6, 1e-6)


def _sd(g, prefix):
    return {k[len(prefix):]: t(g[k]) for k in g.files if k.startswith(prefix)}


def _opts(**kw):
    base = dict(near=2.0, far=6.0, num_coarse=16, num_fine=24, Lx=6, Ld=4, chunksize=1 << 20)
    base.update(kw)
    return O.RenderOptions(**base)


def _check_pipeline(res, g, tag, rtol=2e-4, atol=2e-5):
    for name, v in zip(["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"], res[:6]):
        ref = g[f"{tag}_{name}"]
        close(v.reshape(ref.shape), ref, rtol, atol)
    dex = np.stack([v.numpy().reshape(g[f"{tag}_dex"].shape[1:]) for v in res[6:]], 0)
    ref = g[f"{tag}_dex"]
    # dex depth = z[first sigma > m]; z_fine carries the sample_pdf last-ulp caveat, so compare
    # values tightly rather than bitwise, and require nearly all to be exactly equal
    close(dex, ref, 1e-5, 1e-5)
    assert (dex == ref).mean() > 0.9


def test_pipeline_validation(golden):
    g = golden("pipeline_small")
    mc = lambda x: O.flexible_forward(_sd(g, "coarse."), x, skip_connect_every=2)
    mf = lambda x: O.flexible_forward(_sd(g, "fine."), x, skip_connect_every=2)
    thr = g["thr"].tolist()
    ro, rd = t(g["ro"]), t(g["rd"])
    _check_pipeline(O.render_rays(ro, rd, mc, mf, _opts(), thr), g, "val")
    _check_pipeline(O.render_rays(ro, rd, mc, mf, _opts(white_background=True, lindisp=True), thr), g, "val_wl")


def test_pipeline_train_replayed_rng(golden):
    g = golden("pipeline_small")
    mc = lambda x: O.flexible_forward(_sd(g, "coarse."), x, skip_connect_every=2)
    mf = lambda x: O.flexible_forward(_sd(g, "fine."), x, skip_connect_every=2)
    res = O.render_rays(t(g["ro"]), t(g["rd"]), mc, mf, _opts(perturb=True, noise_std=0.2), g["thr"].tolist(),
                        t_rand=t(g["train_t_rand"]), u=t(g["train_u"]),
                        noise_coarse=t(g["train_noise_c"]), noise_fine=t(g["train_noise_f"]))
    _check_pipeline(res, g, "train")


def test_pipeline_ndc(golden):
    g = golden("pipeline_small")
    mc = lambda x: O.flexible_forward(_sd(g, "ndc_coarse."), x, use_viewdirs=False)
    mf = lambda x: O.flexible_forward(_sd(g, "ndc_fine."), x, use_viewdirs=False)
    H, W = g["HW"]
    o = _opts(near=0.0, far=1.0, no_ndc=False, use_viewdirs=False)
    res = O.render_rays(t(g["ndc_ro"]), t(g["ndc_rd"]), mc, mf, o, g["thr"].tolist(),
                        height=int(H), width=int(W), focal=9.0)
    _check_pipeline(res, g, "ndc")


def test_lego_checkpoint(golden):
    """A trained sigma field (pretrained/lego-lowres) exercises the threshold-depth path."""
    g = golden("lego_lowres")
    H, W = g["HW"]
    ro, rd = O.get_ray_bundle(int(H), int(W), None, t(g["T"]), t(g["K"]))
    mc = lambda x: O.flexible_forward(_sd(g, "coarse."), x)
    mf = lambda x: O.flexible_forward(_sd(g, "fine."), x)
    o = O.RenderOptions(near=2.0, far=6.0, num_coarse=64, num_fine=64, white_background=True, Lx=10, Ld=4)
    res = O.render_rays(ro, rd, mc, mf, o, g["thr"].tolist())
    for name, v in zip(["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"], res[:6]):
        close(v.reshape(g[name].shape), g[name], 5e-4, 5e-5)
    dex = np.stack([v.numpy().reshape(H, W) for v in res[6:]], 0)
    assert (dex == g["dex"]).mean() > 0.98
    assert (g["dex"] > 2.0 + 1e-6).mean() > 0.2          # the thresholds really are crossed


def test_lego_frame(golden):
    """The same checkpoint on the 40x48 view of lego_frame.npz (the flip-rate fixture of the bf16 GPU tests):
    the fp32 oracle reproduces the reference; the oracle under the bf16 operand contract is what the tensor-core
    kernels are held to, so its own distance from the fp32 reference is on record here."""
    g, gw = golden("lego_frame"), golden("lego_lowres")
    H, W = map(int, g["HW"])
    ro, rd = O.get_ray_bundle(H, W, None, t(g["T"]), t(g["K"]))
    o = O.RenderOptions(near=2.0, far=6.0, num_coarse=64, num_fine=64, white_background=True, Lx=10, Ld=4)
    # (operands, max abs, mean abs, same Dex depth value, Dex depth within one fine-sample spacing)
    for bf16, tol_max, tol_mean, same, near_same in ((False, 5e-4, 1e-5, 0.98, 0.999), (True, 0.30, 1.5e-3, 0.70, 0.995)):
        mc = lambda x: O.flexible_forward(_sd(gw, "coarse."), x, bf16=bf16)
        mf = lambda x: O.flexible_forward(_sd(gw, "fine."), x, bf16=bf16)
        res = O.render_rays(ro, rd, mc, mf, o, g["thr"].tolist())
        for name, v in zip(["rgb_c", "acc_c", "rgb_f", "acc_f"], (res[0], res[2], res[3], res[5])):
            d = np.abs(v.numpy().reshape(g[name].shape) - g[name])
            # bf16 operands on a TRAINED field: measured max 0.017 / 0.22 / 0.091 / 0.043 (a few silhouette rays
            # whose alpha sits at a tipping point), mean 3e-4 / 9e-4 / 5e-4 / 1.4e-4 - the contract's price, no GPU
            assert float(d.max()) < tol_max and float(d.mean()) < tol_mean, (bf16, name, float(d.max()), float(d.mean()))
        dex = np.stack([v.numpy().reshape(H, W) for v in res[6:]], 0)
        dd = np.abs(dex - g["dex"])
        assert (dd <= 1e-5).mean() > same and (dd <= 4.0 / 127.0).mean() > near_same, bf16
    assert (g["dex"] > 2.0 + 1e-6).mean() > 0.2


def test_models(golden):
    g = golden("models")
    x = t(g["x90"])
    gen = lambda: torch.Generator().manual_seed(0)
    # functional forwards on the reference-initialised weights are covered through nerf.models
    # in test_host_api.py (needs the package); here: shapes/param counts of the restated layout
    shapes = O.flexible_shapes()
    assert sum(o * i + o for _, o, i in shapes) == int(g["flex8x256_nparams"]) == 595844
    assert [f"{n}.weight" for n, _, _ in shapes] == [k for k in g["flex8x256_keys"].tolist() if k.endswith("weight")]
    sd = O.init_flexible_state_dict(gen())
    assert O.flexible_forward(sd, x).shape == (96, 4)
    a = O.flexible_forward(sd, x)
    b = O.flexible_forward(sd, x, bf16=True)
    assert float((a - b).abs().max()) < 5e-2


def test_tiny(golden):
    g = golden("tiny")
    H, W = g["HW"]
    sd = _sd(g, "model.")
    rgb = O.run_one_iter_of_tinynerf(int(H), int(W), t(g["T"]), t(g["K"]), 2.0, 6.0, 64, 6,
                                     lambda x: O.very_tiny_forward(sd, x))
    close(rgb, g["rgb"], 1e-4, 1e-5)


# ------------------------------------------------------------------ a-10 training iteration
@pytest.mark.parametrize("tag", ["h32", "h128"])
def test_train_gradients_match_reference(golden, tag):
    """Loss, every parameter gradient and one Adam step of the reference's training iteration
    (train_dexnerf_rgb.py:246-281) with its four RNG draws replayed."""
    g = golden("train_grads")
    hidden, layers, skip, white = map(int, g[f"{tag}.cfg"])
    sdc = {k[len(tag) + 8:]: t(v) for k, v in g.items() if k.startswith(f"{tag}.coarse.")}
    sdf = {k[len(tag) + 6:]: t(v) for k, v in g.items() if k.startswith(f"{tag}.fine.")}
    opts = O.RenderOptions(near=2.0, far=6.0, num_coarse=16, num_fine=24, Lx=6, Ld=4, perturb=True,
                           noise_std=0.2, white_background=bool(white))
    loss, cl, fl, gc, gf = O.train_loss_and_grads(
        sdc, sdf, t(g["ro"]), t(g["rd"]), t(g["target"]), opts, g["thr"].tolist(), skip, skip,
        t_rand=t(g[f"{tag}.t_rand"]), u=t(g[f"{tag}.u"]), noise_coarse=t(g[f"{tag}.noise_c"]),
        noise_fine=t(g[f"{tag}.noise_f"]))
    assert abs(float(loss) - float(g[f"{tag}.loss"])) < 2e-6
    assert abs(float(cl) - float(g[f"{tag}.coarse_loss"])) < 2e-6
    for net, grads, sd in (("coarse", gc, sdc), ("fine", gf, sdf)):
        for k, v in grads.items():
            ref = t(g[f"{tag}.grad.{net}.{k}"])
            scale = float(ref.abs().max()) + 1e-12
            assert float((v - ref).abs().max()) <= 2e-4 * scale + 1e-9, (net, k)
        new = O.adam_step(sd, {k: t(g[f"{tag}.grad.{net}.{k}"]) for k in sd}, lr=5e-3)
        for k, v in new.items():
            assert torch.allclose(v, t(g[f"{tag}.adam.{net}.{k}"]), rtol=0, atol=2e-6), (net, k)


# ------------------------------------------------------------------ next rows
def test_depth_error_metrics_and_threshold_selection(golden):
    g = golden("next_rows")
    gt, planes, mask = t(g["metric_gt"]), t(g["metric_planes"]), t(g["metric_mask"])
    for k in range(planes.shape[0]):
        e = O.compute_err_metric(gt, planes[k], mask)
        close([e["depth_abs_err"], e["depth_err2"], e["depth_err4"], e["depth_err8"]], g["metric_errs"][k], 1e-5, 1e-7)
    best, err = O.select_dex_threshold(list(planes), gt)
    assert best == int(g["metric_best"]) and abs(err["depth_abs_err"] - g["metric_errs"][best, 0]) < 1e-4


def test_pose_spherical(golden):
    g = golden("next_rows")
    for args, ref in zip(g["pose_args"], g["poses"]):
        close(O.pose_spherical(*map(float, args)), ref, 1e-6, 1e-6)
