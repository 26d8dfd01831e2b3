"""Generate the golden fixtures in this directory FROM THE REFERENCE ITSELF.

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference package `nerf` from /root/reference/nerf-pytorch with the three
import shims in `_shims/` (torchsearchsorted -> torch.searchsorted, imageio.imread through PIL, empty matplotlib),
runs the hot-path functions of SURVEY.md section 8(a) on seeded CPU inputs and stores inputs and
outputs as small .npz files.  The two 8-layer forwards are the one-line REPAIRS documented in
SURVEY.md section 8(a-3), applied as subclasses of the reference classes so that parameter
shapes, names and init order still come from the reference's own __init__.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/nerf-pytorch"
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)

import nerf as ref  # noqa: E402  (the reference)
import tiny_nerf as ref_tiny  # noqa: E402

torch.set_num_threads(8)


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def rand_pose(g):
    """A world->cam rigid transform with a random rotation and translation."""
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    if torch.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    T = torch.eye(4, dtype=torch.float64)
    T[:3, :3] = q
    T[:3, 3] = torch.randn(3, generator=g, dtype=torch.float64) * 2
    return T.to(torch.float32)


def sigma_field(g, n, S):
    """SURVEY.md section 8(d) operator-level distribution: rgb ~ N(0,1), sigma_raw ~ 30 N(0,1)
    with 5 % spikes U(50, 300) so that every threshold 5..100 is hit."""
    rf = torch.randn(n, S, 4, generator=g)
    sig = 30.0 * torch.randn(n, S, generator=g)
    spikes = torch.rand(n, S, generator=g) < 0.05
    sig = torch.where(spikes, 50 + 250 * torch.rand(n, S, generator=g), sig)
    rf[..., 3] = sig
    return rf


class RepairedFlexible(ref.models.FlexibleNeRFModel):
    """models.py:233-256 with the skip condition __init__ uses (models.py:210)."""

    def forward(self, x):
        xyz, view = x[..., : self.dim_xyz], x[..., self.dim_xyz:]
        x = self.layer1(xyz)
        for i in range(len(self.layers_xyz)):
            if i % self.skip_connect_every == 0 and i > 0:
                x = torch.cat((x, xyz), dim=-1)
            x = self.relu(self.layers_xyz[i](x))
        feat = self.relu(self.fc_feat(x))
        alpha = self.fc_alpha(x)
        x = torch.cat((feat, view), dim=-1)
        for l in self.layers_dir:
            x = self.relu(l(x))
        return torch.cat((self.fc_rgb(x), alpha), dim=-1)


class RepairedPaper(ref.models.PaperNeRFModel):
    """models.py:163-182 with the trunk started from xyz."""

    def forward(self, x):
        xyz, dirs = x[..., : self.dim_xyz], x[..., self.dim_xyz:]
        x = xyz
        for i in range(8):
            if i == 4:
                x = self.layers_xyz[i](torch.cat((xyz, x), -1))
            else:
                x = self.layers_xyz[i](x)
            x = self.relu(x)
        feat = self.fc_feat(x)
        alpha = self.fc_alpha(feat)
        x = self.relu(self.layers_dir[0](torch.cat((feat, dirs), -1)))
        for i in range(1, 3):
            x = self.relu(self.layers_dir[i](x))
        return torch.cat((self.fc_rgb(x), alpha), dim=-1)


def weight_digest(module):
    """Order-dependent fp64 digest of all parameters (pins torch's init order under a seed)."""
    acc, k = 0.0, 1
    for name, p in module.state_dict().items():
        v = p.detach().double().flatten()
        acc += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64)).sum()) * k
        k += 1
    return acc


def make_cfg(num_coarse, num_fine, near, far, perturb, noise_std, white_bg, lindisp, no_ndc=True,
             use_viewdirs=True, chunksize=1 << 20):
    mode = dict(chunksize=chunksize, perturb=perturb, num_coarse=num_coarse, num_fine=num_fine,
                white_background=white_bg, radiance_field_noise_std=noise_std, lindisp=lindisp)
    d = dict(dataset=dict(no_ndc=no_ndc, near=near, far=far),
             nerf=dict(use_viewdirs=use_viewdirs, train=dict(mode, num_random_rays=64),
                       validation=dict(mode)))
    return ref.CfgNode(d)


# ------------------------------------------------------------------ operator-level fixtures
def gen_ops():
    g = torch.Generator().manual_seed(1234)
    out = {}

    # a-1 ray generation (two cameras; one with fx != fy to pin the "/fx for both" quirk)
    for tag, (H, W, fx, fy, cx, cy) in {"a": (6, 8, 11.5, 13.0, 3.7, 2.9),
                                        "b": (5, 3, 50.0, 50.0, 1.5, 2.5)}.items():
        T = rand_pose(g)
        K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
        ro, rd = ref.get_ray_bundle(H, W, None, T, K)
        out[f"ray_{tag}_T"], out[f"ray_{tag}_K"] = npy(T), npy(K)
        out[f"ray_{tag}_HW"] = np.array([H, W])
        out[f"ray_{tag}_ro"], out[f"ray_{tag}_rd"] = npy(ro), npy(rd)

    # ndc_rays
    ro = torch.randn(7, 5, 3, generator=g)
    rd = torch.randn(7, 5, 3, generator=g)
    rd[..., 2] = -rd[..., 2].abs() - 0.5
    o, d = ref.ndc_rays(7, 5, 9.25, 1.0, ro, rd)
    out.update(ndc_ro=npy(ro), ndc_rd=npy(rd), ndc_o=npy(o), ndc_d=npy(d),
               ndc_args=np.array([7, 5, 9.25, 1.0]))

    # a-2 positional encoding
    x = torch.randn(50, 3, generator=g) * 3.0
    out["pe_x"] = npy(x)
    out["pe_L10"] = npy(ref.positional_encoding(x, 10, True, True))
    out["pe_L4"] = npy(ref.positional_encoding(x, 4, True, True))
    out["pe_L6_lin"] = npy(ref.positional_encoding(x, 6, True, False))
    out["pe_L5_noinput"] = npy(ref.positional_encoding(x, 5, False, True))
    out["pe_L0"] = npy(ref.positional_encoding(x, 0, True, True))
    out["pe_kat_in"] = np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    out["pe_kat_out"] = npy(ref.positional_encoding(torch.tensor([[1.0, 2.0, 3.0]]), 2))

    # a-5 cumprod_exclusive
    c = torch.rand(5, 7, generator=g)
    out["cp_in"], out["cp_out"] = npy(c), npy(ref.cumprod_exclusive(c))
    out["cp_kat"] = npy(ref.cumprod_exclusive(torch.tensor([[0.5, 0.5, 0.5, 0.5]])))

    # a-6 compositing + Dex depth
    thr = [float(m) for m in np.arange(5, 105, 5)]
    out["vr_thr"] = np.array(thr, dtype=np.float32)
    for tag, (n, S, white, std) in {"s48": (37, 48, False, 0.0), "s64w": (33, 64, True, 0.0),
                                    "s192n": (9, 192, False, 0.2), "s5": (4, 5, False, 0.0)}.items():
        rf = sigma_field(g, n, S)
        if tag == "s48":
            rf[0, :, 3] = -1.0          # nothing absorbs: acc 0, disp NaN, dex = z[0]
            rf[1, :, 3] = 0.0
            rf[2, :-1, 3] = -5.0        # only the last sample (1e10 dist) absorbs
            rf[2, -1, 3] = 0.5
        z = torch.sort(0.3 + 3.7 * torch.rand(n, S, generator=g), dim=-1).values
        rd = torch.randn(n, 3, generator=g)
        noise = None
        if std > 0:
            torch.manual_seed(77)
            noise = torch.randn(n, S) * std
            torch.manual_seed(77)
        res = ref.volume_render_radiance_field(rf, z, rd, std, white, thr)
        out[f"vr_{tag}_rf"], out[f"vr_{tag}_z"], out[f"vr_{tag}_rd"] = npy(rf), npy(z), npy(rd)
        out[f"vr_{tag}_white"] = np.array(white)
        if noise is not None:
            out[f"vr_{tag}_noise"] = npy(noise)
        for name, v in zip(["rgb", "disp", "acc", "weights", "depth"], res[:5]):
            out[f"vr_{tag}_{name}"] = npy(v)
        out[f"vr_{tag}_dex"] = np.stack([npy(v) for v in res[5:]], 0)
    # known-answer rows
    sig = torch.tensor([[0, 1, 20, 3, 30, 0], [1, 1, 1, 1, 1, 1], [16, 0, 0, 0, 0, 0]], dtype=torch.float32)
    rf = torch.zeros(3, 6, 4)
    rf[..., 3] = sig
    z = torch.linspace(1, 2, 6).expand(3, 6).contiguous()
    res = ref.volume_render_radiance_field(rf, z, torch.ones(3, 3), 0.0, False, [5.0, 10.0, 15.0])
    out["vr_kat_dex"] = np.stack([npy(v) for v in res[5:]], 0)

    # a-7 sample_pdf (nerf.sample_pdf is sample_pdf_2; v1 kept to pin the equivalence)
    for tag, (n, B, Nf) in {"c2": (40, 63, 128), "c5": (6, 127, 256), "odd": (5, 9, 17)}.items():
        bins = torch.sort(2 + 4 * torch.rand(n, B, generator=g), dim=-1).values
        w = torch.rand(n, B - 1, generator=g) ** 8
        w[0] = 0.0
        w[1, : (B - 1) // 2] = 0.0
        det = ref.sample_pdf(bins, w, Nf, det=True)
        det_v1 = ref.nerf_helpers.sample_pdf(bins, w, Nf, det=True)
        torch.manual_seed(5)
        u = torch.rand(n, Nf)
        torch.manual_seed(5)
        rnd = ref.sample_pdf(bins, w, Nf, det=False)
        out[f"sp_{tag}_bins"], out[f"sp_{tag}_w"] = npy(bins), npy(w)
        out[f"sp_{tag}_det"], out[f"sp_{tag}_det_v1"] = npy(det), npy(det_v1)
        out[f"sp_{tag}_u"], out[f"sp_{tag}_rnd"] = npy(u), npy(rnd)
        # the (cdf, u) -> inds boundary, as the reference computes it
        wp = w + 1e-5
        cdf = torch.cumsum(wp / torch.sum(wp, -1, keepdim=True), -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        out[f"sp_{tag}_cdf"] = npy(cdf)
        out[f"sp_{tag}_inds_u"] = npy(torch.searchsorted(cdf, u, right=True))
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("ops.npz", len(out), "arrays")


# ------------------------------------------------------------------ model fixtures
def gen_models():
    out = {}
    g = torch.Generator().manual_seed(99)
    x90 = torch.cat([ref.positional_encoding(torch.randn(96, 3, generator=g) * 2, 10),
                     ref.positional_encoding(torch.nn.functional.normalize(torch.randn(96, 3, generator=g)), 4)], -1)
    out["x90"] = npy(x90)
    torch.manual_seed(42)
    m = RepairedFlexible(num_layers=8, hidden_size=256, skip_connect_every=4,
                         num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    out["flex8x256_digest"] = np.array(weight_digest(m))
    out["flex8x256_nparams"] = np.array(sum(p.numel() for p in m.parameters()))
    out["flex8x256_keys"] = np.array(list(m.state_dict().keys()))
    out["flex8x256_out"] = npy(m(x90))
    torch.manual_seed(42)
    m = RepairedFlexible(num_layers=8, hidden_size=128, skip_connect_every=3,
                         num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    out["flex8x128s3_digest"] = np.array(weight_digest(m))
    out["flex8x128s3_out"] = npy(m(x90))
    torch.manual_seed(42)
    m = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)  # as-run 4x128
    out["flex4x128_digest"] = np.array(weight_digest(m))
    out["flex4x128_out"] = npy(m(x90))
    torch.manual_seed(42)
    m = RepairedPaper(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    out["paper_digest"] = np.array(weight_digest(m))
    out["paper_nparams"] = np.array(sum(p.numel() for p in m.parameters()))
    out["paper_keys"] = np.array(list(m.state_dict().keys()))
    out["paper_out"] = npy(m(x90))
    torch.manual_seed(42)
    m = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=6, num_encoding_fn_dir=4, use_viewdirs=False)
    out["flex_noview_digest"] = np.array(weight_digest(m))
    out["flex_noview_out"] = npy(m(x90[:, :39]))
    np.savez_compressed(os.path.join(HERE, "models.npz"), **out)
    print("models.npz")


# ------------------------------------------------------------------ pipeline fixtures
def boosted_small_model(seed):
    """A small FlexibleNeRFModel whose sigma head is scaled so that densities span 0..200."""
    torch.manual_seed(seed)
    m = RepairedFlexible(num_layers=5, hidden_size=32, skip_connect_every=2,
                         num_encoding_fn_xyz=6, num_encoding_fn_dir=4)
    with torch.no_grad():
        m.fc_alpha.weight.mul_(400.0)
        m.fc_alpha.bias.fill_(5.0)
    return m


def gen_pipeline():
    out = {}
    g = torch.Generator().manual_seed(7)
    H, W = 6, 8
    T = rand_pose(g)
    T[:3, 3] = torch.tensor([0.1, -0.2, 3.0])
    K = torch.tensor([[9.0, 0, 4.0], [0, 9.0, 3.0], [0, 0, 1]])
    ro, rd = ref.get_ray_bundle(H, W, None, T, K)
    mc, mf = boosted_small_model(11), boosted_small_model(12)
    for k, v in mc.state_dict().items():
        out["coarse." + k] = npy(v)
    for k, v in mf.state_dict().items():
        out["fine." + k] = npy(v)
    out.update(T=npy(T), K=npy(K), HW=np.array([H, W]), ro=npy(ro), rd=npy(rd))
    thr = [5.0, 20.0, 60.0, 100.0]
    out["thr"] = np.array(thr, dtype=np.float32)
    enc_x = ref.get_embedding_function(6, True, True)
    enc_d = ref.get_embedding_function(4, True, True)
    names = ["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"]

    # validation mode (deterministic), plain and white-background / lindisp variants
    for tag, kw in {"val": dict(white_bg=False, lindisp=False),
                    "val_wl": dict(white_bg=True, lindisp=True)}.items():
        cfg = make_cfg(16, 24, 2.0, 6.0, False, 0.0, **kw)
        with torch.no_grad():
            res = ref.run_one_iter_of_nerf(H, W, 9.0, mc, mf, ro, rd, cfg, mode="validation",
                                           encode_position_fn=enc_x, encode_direction_fn=enc_d,
                                           m_thres_cand=thr)
        for nme, v in zip(names, res[:6]):
            out[f"{tag}_{nme}"] = npy(v)
        out[f"{tag}_dex"] = np.stack([npy(v) for v in res[6:]], 0)

    # train mode: replay the reference's four RNG draws (train_utils.py:132,
    # volume_rendering_utils.py:32-39, nerf_helpers.py:279-283, then the fine noise)
    cfg = make_cfg(16, 24, 2.0, 6.0, True, 0.2, False, False)
    n = H * W
    torch.manual_seed(2024)
    t_rand = torch.rand(n, 16)
    noise_c = torch.randn(n, 16) * 0.2
    u = torch.rand(n, 24)
    noise_f = torch.randn(n, 40) * 0.2
    torch.manual_seed(2024)
    with torch.no_grad():
        res = ref.run_one_iter_of_nerf(H, W, 9.0, mc, mf, ro.reshape(-1, 3), rd.reshape(-1, 3), cfg,
                                       mode="train", encode_position_fn=enc_x,
                                       encode_direction_fn=enc_d, m_thres_cand=thr)
    out.update(train_t_rand=npy(t_rand), train_noise_c=npy(noise_c), train_u=npy(u),
               train_noise_f=npy(noise_f))
    for nme, v in zip(names, res[:6]):
        out[f"train_{nme}"] = npy(v)
    out["train_dex"] = np.stack([npy(v) for v in res[6:]], 0)

    # NDC variant (no_ndc False), no view dirs in the nets
    torch.manual_seed(21)
    mc2 = ref.models.FlexibleNeRFModel(num_layers=4, hidden_size=32, num_encoding_fn_xyz=6,
                                       num_encoding_fn_dir=4, use_viewdirs=False)
    mf2 = ref.models.FlexibleNeRFModel(num_layers=4, hidden_size=32, num_encoding_fn_xyz=6,
                                       num_encoding_fn_dir=4, use_viewdirs=False)
    for k, v in mc2.state_dict().items():
        out["ndc_coarse." + k] = npy(v)
    for k, v in mf2.state_dict().items():
        out["ndc_fine." + k] = npy(v)
    ro2 = torch.randn(H, W, 3, generator=g) * 0.1
    rd2 = torch.randn(H, W, 3, generator=g) * 0.3
    rd2[..., 2] = -1.0
    cfg = make_cfg(16, 24, 0.0, 1.0, False, 0.0, False, False, no_ndc=False, use_viewdirs=False)
    with torch.no_grad():
        res = ref.run_one_iter_of_nerf(H, W, 9.0, mc2, mf2, ro2, rd2, cfg, mode="validation",
                                       encode_position_fn=enc_x, encode_direction_fn=None,
                                       m_thres_cand=thr)
    out.update(ndc_ro=npy(ro2), ndc_rd=npy(rd2))
    for nme, v in zip(names, res[:6]):
        out[f"ndc_{nme}"] = npy(v)
    out["ndc_dex"] = np.stack([npy(v) for v in res[6:]], 0)
    np.savez_compressed(os.path.join(HERE, "pipeline_small.npz"), **out)
    print("pipeline_small.npz")


def gen_lego():
    """A trained sigma field: pretrained/lego-lowres (4x128 Flexible, Lx=10, Ld=4), a 10x12 view."""
    ck = torch.load(os.path.join(REF, "pretrained/lego-lowres/checkpoint199999.ckpt"),
                    map_location="cpu", weights_only=False)
    mc = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    mf = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    mc.load_state_dict(ck["model_coarse_state_dict"])
    mf.load_state_dict(ck["model_fine_state_dict"])
    out = {}
    for k, v in mc.state_dict().items():
        out["coarse." + k] = npy(v).astype(np.float32)
    for k, v in mf.state_dict().items():
        out["fine." + k] = npy(v).astype(np.float32)
    H, W = 10, 12
    # blender "lego" style camera on the r=4 sphere, expressed in the fork's world->cam form;
    # the focal is scaled so that the 10x12 crop still sees the object.
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.nerf_oracle import pose_spherical_world2cam
    T = pose_spherical_world2cam(30.0, -30.0, 4.0)
    # upstream lego checkpoints were trained with OpenGL rays (-z forward, y up); the fork's
    # get_ray_bundle emits OpenCV rays, which our pose helper already compensates for.
    K = torch.tensor([[14.0, 0, 6.0], [0, 14.0, 5.0], [0, 0, 1]])
    ro, rd = ref.get_ray_bundle(H, W, None, T, K)
    thr = [float(m) for m in np.arange(5, 105, 5)]
    cfg = make_cfg(64, 64, 2.0, 6.0, False, 0.0, True, False)
    with torch.no_grad():
        res = ref.run_one_iter_of_nerf(H, W, 14.0, mc, mf, ro, rd, cfg, mode="validation",
                                       encode_position_fn=ref.get_embedding_function(10, True, True),
                                       encode_direction_fn=ref.get_embedding_function(4, True, True),
                                       m_thres_cand=thr)
    out.update(T=npy(T), K=npy(K), HW=np.array([H, W]), thr=np.array(thr, dtype=np.float32))
    for nme, v in zip(["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"], res[:6]):
        out[nme] = npy(v)
    out["dex"] = np.stack([npy(v) for v in res[6:]], 0)
    print("lego acc_f mean", float(res[5].mean()), "dex>near frac",
          float((res[6] > 2.0 + 1e-6).float().mean()), float((res[-1] > 2.0 + 1e-6).float().mean()))
    np.savez_compressed(os.path.join(HERE, "lego_lowres.npz"), **out)
    print("lego_lowres.npz")


def gen_lego_frame():
    """The same trained lego-lowres checkpoint on a 40x48 view (1 920 rays x 20 thresholds): enough rays for the
    Dex-depth flip-rate table of the bf16 tensor-core path.  Weights are NOT stored again
    (lego_lowres.npz holds them); outputs only."""
    ck = torch.load(os.path.join(REF, "pretrained/lego-lowres/checkpoint199999.ckpt"),
                    map_location="cpu", weights_only=False)
    mc = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    mf = ref.models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    mc.load_state_dict(ck["model_coarse_state_dict"])
    mf.load_state_dict(ck["model_fine_state_dict"])
    H, W = 40, 48
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.nerf_oracle import pose_spherical_world2cam
    T = pose_spherical_world2cam(30.0, -30.0, 4.0)
    K = torch.tensor([[56.0, 0, 24.0], [0, 56.0, 20.0], [0, 0, 1]])
    ro, rd = ref.get_ray_bundle(H, W, None, T, K)
    thr = [float(m) for m in np.arange(5, 105, 5)]
    cfg = make_cfg(64, 64, 2.0, 6.0, False, 0.0, True, False)
    with torch.no_grad():
        res = ref.run_one_iter_of_nerf(H, W, 56.0, mc, mf, ro, rd, cfg, mode="validation",
                                       encode_position_fn=ref.get_embedding_function(10, True, True),
                                       encode_direction_fn=ref.get_embedding_function(4, True, True),
                                       m_thres_cand=thr)
    out = dict(T=npy(T), K=npy(K), HW=np.array([H, W]), thr=np.array(thr, dtype=np.float32))
    for nme, v in zip(["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"], res[:6]):
        out[nme] = npy(v)
    out["dex"] = np.stack([npy(v) for v in res[6:]], 0)
    print("lego frame acc_f mean", float(res[5].mean()), "dex>near frac",
          float((res[6] > 2.0 + 1e-6).float().mean()), float((res[-1] > 2.0 + 1e-6).float().mean()))
    np.savez_compressed(os.path.join(HERE, "lego_frame.npz"), **out)
    print("lego_frame.npz")


def gen_tiny():
    """BASELINE config 1 shrunk to 20x20: tiny_nerf.py functions with a 5-arg get_ray_bundle."""
    H = W = 20
    K = torch.tensor([[27.6, 0, 10.0], [0, 27.6, 10.0], [0, 0, 1]])
    T = torch.eye(4)
    T[2, 3] = 4.0
    torch.manual_seed(9458)
    model = ref_tiny.VeryTinyNerfModel(num_encoding_functions=6)
    ro, rd = ref.get_ray_bundle(H, W, None, T, K)
    pts, z = ref_tiny.compute_query_points_from_rays(ro, rd, 2.0, 6.0, 64, randomize=False)
    enc = ref.positional_encoding(pts.reshape(-1, 3), 6)
    with torch.no_grad():
        rf = model(enc).reshape(H, W, 64, 4)
        rgb, depth, acc = ref_tiny.render_volume_density(rf, ro, z)
    out = dict(T=npy(T), K=npy(K), HW=np.array([H, W]), rgb=npy(rgb), depth=npy(depth), acc=npy(acc),
               digest=np.array(weight_digest(model)))
    for k, v in model.state_dict().items():
        out["model." + k] = npy(v)
    np.savez_compressed(os.path.join(HERE, "tiny.npz"), **out)
    print("tiny.npz")


def gen_next_rows():
    """compute_err_metric (train_utils.py:9-30) on synthetic depth planes
    with the reference's mask rule, the threshold selection loop of train_dexnerf_rgb.py:392-404
    restated around it, and pose_spherical (load_blender.py:33-38)."""
    from nerf.load_blender import pose_spherical
    out = {}
    g = torch.Generator().manual_seed(808)
    H, W, T = 27, 48, 20
    gt = 0.3 + 1.2 * torch.rand(H, W, generator=g)
    gt[torch.rand(H, W, generator=g) < 0.1] = 0.0                     # holes in the depth sensor
    planes = gt[None] + 0.02 * torch.randn(T, H, W, generator=g) * torch.linspace(2.0, 0.1, T)[:, None, None].abs()
    planes[5] = gt + 0.0005 * torch.randn(H, W, generator=g)          # the best candidate
    mask = (gt > 0) & (gt < 1.25)
    errs = []
    for k in range(T):
        e = ref.compute_err_metric(gt, planes[k], mask)
        errs.append([e["depth_abs_err"], e["depth_err2"], e["depth_err4"], e["depth_err8"]])
    errs = np.array(errs, dtype=np.float64)
    out.update(metric_gt=npy(gt), metric_planes=npy(planes), metric_mask=npy(mask), metric_errs=errs,
               metric_best=np.array(int(np.argmin(errs[:, 0]))))
    out["poses"] = np.stack([np.asarray(pose_spherical(a, p, r), dtype=np.float64)
                             for a, p, r in [(30.0, -30.0, 4.0), (-180.0, -30.0, 4.0), (99.0, 10.0, 2.5)]], 0)
    out["pose_args"] = np.array([(30.0, -30.0, 4.0), (-180.0, -30.0, 4.0), (99.0, 10.0, 2.5)])
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
    print("next_rows.npz best", int(out["metric_best"]), errs[int(out["metric_best"])])


def gen_datasets():
    """the three dataset loaders of the reference on the tiny synthetic
    datasets of tests/dataset_fixture.py (files are rebuilt byte-identically by the test)."""
    import tempfile
    sys.path.insert(0, os.path.dirname(HERE))
    import dataset_fixture as DF
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        b = DF.build_blender(os.path.join(tmp, "blender"))
        for tag, kw in (("full", dict()), ("half", dict(half_res=True, testskip=2)), ("debug", dict(debug=True))):
            imgs, poses, render_poses, hwf, i_split = ref.load_blender_data(b, **kw)
            out.update({f"bl_{tag}_imgs": npy(imgs), f"bl_{tag}_poses": npy(poses), f"bl_{tag}_render": npy(render_poses),
                        f"bl_{tag}_hwf": np.asarray(hwf, dtype=np.float64),
                        f"bl_{tag}_split": np.concatenate([np.asarray([len(x) for x in i_split])] + list(i_split))})
        m = DF.build_messytable(os.path.join(tmp, "messy"))
        for tag, kw in (("ir", dict(half_res=True)), ("full", dict(half_res=False)),
                        ("rgb", dict(half_res=True, imgname="rgb.png", is_real_rgb=True)), ("debug", dict(debug=True))):
            imgs, poses, render_poses, hwf, i_split, intr, depths = ref.load_messytable_data(m, **kw)
            out.update({f"mt_{tag}_imgs": npy(imgs), f"mt_{tag}_poses": npy(poses), f"mt_{tag}_render": npy(render_poses),
                        f"mt_{tag}_hwf": np.asarray(hwf, dtype=np.float64), f"mt_{tag}_intr": npy(intr),
                        f"mt_{tag}_depths": npy(depths),
                        f"mt_{tag}_split": np.concatenate([np.asarray([len(x) for x in i_split])] + list(i_split))})
        ll = DF.build_llff(os.path.join(tmp, "llff"))
        for tag, kw in (("std", dict(factor=8)), ("sph", dict(factor=8, spherify=True)),
                        ("zflat", dict(factor=8, path_zflat=True, recenter=False))):
            if tag == "zflat":
                continue   # the reference divides N_views by 2 into a float and np.linspace then raises (load_llff.py:319,330)
            images, poses, bds, render_poses, i_test = ref.load_llff_data(ll, **kw)
            out.update({f"ll_{tag}_imgs": npy(images), f"ll_{tag}_poses": npy(poses), f"ll_{tag}_bds": npy(bds),
                        f"ll_{tag}_render": npy(render_poses), f"ll_{tag}_itest": np.asarray(i_test)})
    # depth_error_img (train_utils.py:45-70) on a seeded pair of depth maps
    g = torch.Generator().manual_seed(11)
    gt = 300 + 3000 * torch.rand(1, 40, 260, generator=g)
    est = gt + torch.randn(1, 40, 260, generator=g) * torch.tensor([0.1, 1.0, 10.0, 100.0]).repeat(65)[None, None, :]
    mask = torch.rand(1, 40, 260, generator=g) > 0.2
    out["dei_gt"], out["dei_est"], out["dei_mask"] = npy(gt), npy(est), npy(mask)
    out["dei_img"] = ref.depth_error_img(est, gt, mask)
    out["dei_img_thr"] = ref.depth_error_img(est, gt, mask, abs_thres=2.5)
    np.savez_compressed(os.path.join(HERE, "datasets.npz"), **out)
    print("datasets.npz", len(out), "arrays")


def gen_train_grads():
    """One TRAINING iteration of the reference (train_dexnerf_rgb.py:246-278): train-mode
    run_one_iter_of_nerf with the four RNG draws replayed, loss = mse(rgb_coarse, target) +
    mse(rgb_fine, target), loss.backward().  Stores the loss and every parameter gradient of both
    nets, plus the Adam-updated parameters after one optimizer.step() (lr 5e-3)."""
    out = {}
    g = torch.Generator().manual_seed(17)
    H, W = 6, 8
    T = rand_pose(g)
    T[:3, 3] = torch.tensor([0.1, -0.2, 3.0])
    K = torch.tensor([[9.0, 0, 4.0], [0, 9.0, 3.0], [0, 0, 1]])
    ro, rd = ref.get_ray_bundle(H, W, None, T, K)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    n = ro.shape[0]
    target = torch.rand(n, 3, generator=g)
    thr = [5.0, 20.0, 60.0, 100.0]
    enc_x = ref.get_embedding_function(6, True, True)
    enc_d = ref.get_embedding_function(4, True, True)
    for tag, (hidden, layers, skip, white) in {"h32": (32, 5, 2, False), "h128": (128, 8, 3, True)}.items():
        nets = []
        for seed in (31, 32):
            torch.manual_seed(seed)
            m = RepairedFlexible(num_layers=layers, hidden_size=hidden, skip_connect_every=skip,
                                 num_encoding_fn_xyz=6, num_encoding_fn_dir=4)
            with torch.no_grad():
                m.fc_alpha.weight.mul_(20.0)
                m.fc_alpha.bias.fill_(1.0)
            nets.append(m)
        mc, mf = nets
        for k, v in mc.state_dict().items():
            out[f"{tag}.coarse." + k] = npy(v).copy()
        for k, v in mf.state_dict().items():
            out[f"{tag}.fine." + k] = npy(v).copy()
        cfg = make_cfg(16, 24, 2.0, 6.0, True, 0.2, white, False)
        torch.manual_seed(4040)
        t_rand = torch.rand(n, 16)
        noise_c = torch.randn(n, 16) * 0.2
        u = torch.rand(n, 24)
        noise_f = torch.randn(n, 40) * 0.2
        opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=5e-3)
        torch.manual_seed(4040)
        res = ref.run_one_iter_of_nerf(H, W, 9.0, mc, mf, ro, rd, cfg, mode="train", encode_position_fn=enc_x,
                                       encode_direction_fn=enc_d, m_thres_cand=thr)
        coarse_loss = torch.nn.functional.mse_loss(res[0][..., :3], target[..., :3])
        fine_loss = torch.nn.functional.mse_loss(res[3][..., :3], target[..., :3])
        loss = coarse_loss + fine_loss
        loss.backward()
        out.update({f"{tag}.t_rand": npy(t_rand), f"{tag}.noise_c": npy(noise_c), f"{tag}.u": npy(u),
                    f"{tag}.noise_f": npy(noise_f), f"{tag}.loss": npy(loss),
                    f"{tag}.coarse_loss": npy(coarse_loss), f"{tag}.fine_loss": npy(fine_loss),
                    f"{tag}.cfg": np.array([hidden, layers, skip, int(white)])})
        for k, p in mc.named_parameters():
            out[f"{tag}.grad.coarse." + k] = npy(p.grad).copy()
        for k, p in mf.named_parameters():
            out[f"{tag}.grad.fine." + k] = npy(p.grad).copy()
        opt.step()
        for k, p in mc.named_parameters():
            out[f"{tag}.adam.coarse." + k] = npy(p).copy()
        for k, p in mf.named_parameters():
            out[f"{tag}.adam.fine." + k] = npy(p).copy()
        print(tag, "loss", float(loss), "grad norm coarse",
              float(sum((p.grad ** 2).sum() for p in mc.parameters()) ** 0.5))
    out.update(T=npy(T), K=npy(K), HW=np.array([H, W]), ro=npy(ro), rd=npy(rd), target=npy(target),
               thr=np.array(thr, dtype=np.float32))
    np.savez_compressed(os.path.join(HERE, "train_grads.npz"), **out)
    print("train_grads.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1:          # regenerate selected fixtures only: make_golden.py train_grads ...
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_ops()
    gen_models()
    gen_pipeline()
    gen_lego()
    gen_lego_frame()
    gen_tiny()
    gen_train_grads()
    gen_next_rows()
    gen_datasets()
