"""End-to-end parity of the path that is BENCHMARKED - the tcgen05 tensor-core MLP (bf16 operands, fp32
accumulate) - on the configurations the fp32 file pins: the trained lego-lowres checkpoint (the only realistic
sigma field the reference ships; 27 % of its rays cross the Dex thresholds), BASELINE config 5 (128 + 256
samples) and config 2.  Two kinds of bars, both written out below:

  * against the CPU oracle under the SAME operand contract (oracle.flexible_forward(bf16=True)): the kernels'
    own error - accumulation order only;
  * against the reference's fp32 outputs (golden fixtures made from the unmodified reference): the cost of the
    bf16 operand contract itself.  On random-init / boosted synthetic fields that is <= 2e-3 max abs on rgb / acc
    (BASELINE.json's tolerance).  On the TRAINED field it is not: the CPU oracle with bf16 operands - no GPU
    involved - is off by up to 0.09 (rgb_fine) / 0.22 (acc_coarse) on a handful of silhouette rays whose alpha sits
    at a tipping point, while the MEAN error is 5e-4 (tests/test_oracle_golden.py::test_lego_frame records it).
    The bars for the trained field are therefore: mean and 99th percentile of the maps, Dex depth within ONE
    sample spacing for >= 99.5 % of (ray, threshold) pairs, and the per-threshold flip-rate table is printed and
    written to gpurun_out/bf16_parity.json (DESIGN.md section 2 quotes it).
"""
import json
import os

import numpy as np
import pytest
import torch

import nerf
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
t = torch.from_numpy
NAMES = ["rgb_c", "depth_c", "acc_c", "rgb_f", "depth_f", "acc_f"]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_cfg(nc, nf, near, far, white=False):
    mode = dict(chunksize=1 << 20, perturb=False, num_coarse=nc, num_fine=nf, white_background=white,
                radiance_field_noise_std=0.0, lindisp=False)
    return nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=near, far=far),
                             nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))


def _record(key, value):
    """Append a measurement to gpurun_out/bf16_parity.json (scratch; copied into profiles/ by hand)."""
    path = os.path.join(ROOT, "gpurun_out", "bf16_parity.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[key] = value
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def _lego(golden):
    gw = golden("lego_lowres")
    sds = [{k[len(p):]: t(gw[k]) for k in gw.files if k.startswith(p)} for p in ("coarse.", "fine.")]
    nets = []
    for sd in sds:
        m = nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
        m.load_state_dict(sd)
        nets.append(m.cuda())
    return nets, sds


def _render_lego(nets, g):
    H, W = map(int, g["HW"])
    focal = float(g["K"][0, 0])
    with torch.no_grad():
        ro, rd = nerf.get_ray_bundle(H, W, None, t(g["T"]).cuda(), t(g["K"]).cuda())
        return nerf.run_one_iter_of_nerf(H, W, focal, nets[0], nets[1], ro, rd, make_cfg(64, 64, 2.0, 6.0, white=True),
                                         mode="validation",
                                         encode_position_fn=nerf.get_embedding_function(10, True, True),
                                         encode_direction_fn=nerf.get_embedding_function(4, True, True),
                                         m_thres_cand=g["thr"].tolist())


def _stats(res, ref_maps, ref_dex, spacing):
    out = {}
    for name, v in zip(NAMES, res[:6]):
        d = np.abs(v.cpu().numpy().reshape(ref_maps[name].shape) - ref_maps[name])
        out[name] = dict(max=float(d.max()), mean=float(d.mean()), p99=float(np.quantile(d, 0.99)))
    dex = torch.stack(res[6:], 0).cpu().numpy().reshape(ref_dex.shape)
    dd = np.abs(dex - ref_dex).reshape(ref_dex.shape[0], -1)
    out["dex_same_sample"] = float((dd <= 1e-5).mean())                       # the very same depth value
    out["dex_within_one_spacing"] = float((dd <= spacing).mean())
    out["dex_flip_rate_per_threshold"] = [round(float(x), 4) for x in (dd > 1e-5).mean(1)]       # depth value differs at all
    out["dex_off_by_more_than_one_spacing_per_threshold"] = [round(float(x), 4) for x in (dd > spacing).mean(1)]
    out["dex_max_abs"] = float(dd.max())
    return out


@pytest.mark.parametrize("fixture", ["lego_lowres", "lego_frame"])
def test_lego_checkpoint_on_the_tensor_core_path(golden, fixture):
    """pretrained/lego-lowres (4 x 128, what every reference script instantiates), validation render, 64 + 64
    samples, white background, T = 20 - the fp32 file's test_lego_checkpoint on the tensor cores."""
    assert nerf.get_precision() == "bf16"
    g = golden(fixture)
    nets, sds = _lego(golden)
    res = _render_lego(nets, g)
    H, W = map(int, g["HW"])
    assert len(res) == 26 and res[3].shape == (H, W, 3)
    spacing = (6.0 - 2.0) / 127.0          # mean distance of the 128 fine samples; a coarse bin is twice that
    ref_maps = {k: g[k] for k in NAMES}
    s_ref = _stats(res, ref_maps, g["dex"], spacing)
    # the CPU oracle under the same bf16 operand contract
    ro, rd = O.get_ray_bundle(H, W, None, t(g["T"]), t(g["K"]))
    o = O.RenderOptions(near=2.0, far=6.0, num_coarse=64, num_fine=64, white_background=True, Lx=10, Ld=4)
    orc = O.render_rays(ro, rd, lambda x: O.flexible_forward(sds[0], x, bf16=True),
                        lambda x: O.flexible_forward(sds[1], x, bf16=True), o, g["thr"].tolist())
    omaps = {k: v.numpy().reshape(g[k].shape) for k, v in zip(NAMES, orc[:6])}
    odex = np.stack([v.numpy().reshape(H, W) for v in orc[6:]], 0)
    s_orc = _stats(res, omaps, odex, spacing)
    print("\n[%s] tensor-core path vs the reference (fp32):" % fixture, json.dumps(s_ref))
    print("[%s] tensor-core path vs the bf16-contract oracle:" % fixture, json.dumps(s_orc))
    _record(fixture, dict(vs_reference_fp32=s_ref, vs_bf16_oracle=s_orc, rays=H * W,
                          crossing_fraction=float((g["dex"] > 2.0 + 1e-6).mean())))
    # ---- bars against the reference (the price of bf16 operands on a trained field; see the module docstring)
    for name in ("rgb_c", "rgb_f", "acc_c", "acc_f"):
        assert s_ref[name]["mean"] < 2e-3, (name, s_ref[name])            # north-star figure holds for the MEAN
        assert s_ref[name]["max"] < 0.30, (name, s_ref[name])
    if H * W >= 1000:       # percentiles need rays: the 120-ray fixture's 99th percentile IS its second-worst ray
        assert s_ref["rgb_f"]["p99"] < 2.5e-2 and s_ref["acc_f"]["p99"] < 1e-2
    assert s_ref["depth_f"]["mean"] < 0.5 * spacing                        # expected depth: within a sample spacing on average
    assert s_ref["dex_within_one_spacing"] > 0.995                          # Dex depth: same or neighbouring sample
    assert max(s_ref["dex_off_by_more_than_one_spacing_per_threshold"]) < 0.02
    # ---- bars against the oracle under the same contract: accumulation order only; the same tipping-point rays
    # amplify it, so the maps are held to the mean / p99 and the Dex depths to the sample spacing
    # (measured: 10x12 view max 7.7e-4; 40x48 view max 5.8e-3 on one ray, p99 3.3e-4, mean 2e-5, 99.3 % of the Dex
    # depths the very same value)
    for name in ("rgb_c", "rgb_f", "acc_c", "acc_f"):
        assert s_orc[name]["mean"] < 1e-4 and s_orc[name]["p99"] < 2e-3 and s_orc[name]["max"] < 1.5e-2, (name, s_orc[name])
    assert s_orc["dex_within_one_spacing"] > 0.999 and s_orc["dex_same_sample"] > 0.97


def test_c5_128_256_on_the_tensor_core_path():
    """BASELINE config 5 (S_fine = 384, near 0.3 / far 4, T = 20) - what `bench.py --workload c5` times - on the
    tensor-core path against the oracle: same operand contract <= 2e-3 rgb / acc max abs (the boosted field is
    sharp, so a few Dex depths move to the neighbouring sample); fp32 oracle: BASELINE's 2e-3 on rgb / acc."""
    assert nerf.get_precision() == "bf16"
    torch.manual_seed(5)
    mc, mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4), nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    with torch.no_grad():
        for m in (mc, mf):
            m.fc_alpha.weight.mul_(200.0)
            m.fc_alpha.bias.fill_(2.0)
    sdc = {k: v.detach().clone() for k, v in mc.state_dict().items()}
    sdf = {k: v.detach().clone() for k, v in mf.state_dict().items()}
    mc, mf = mc.cuda(), mf.cuda()
    K = torch.tensor([[900.0, 0, 640.0], [0, 900.0, 360.0], [0, 0, 1]])
    T = torch.eye(4)
    T[2, 3] = 1.2
    thr = [float(m) for m in range(5, 105, 5)]
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    with torch.no_grad():
        ro, rd = nerf.get_ray_bundle(720, 1280, None, T.cuda(), K.cuda(), row_start=354, row_count=12)
        ro, rd = ro[:, 632:648].contiguous(), rd[:, 632:648].contiguous()
        res = nerf.run_one_iter_of_nerf(720, 1280, 900.0, mc, mf, ro, rd, make_cfg(128, 256, 0.3, 4.0), mode="validation",
                                        encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=thr)
    assert len(res) == 26 and res[3].shape == (12, 16, 3)
    opts = O.RenderOptions(near=0.3, far=4.0, num_coarse=128, num_fine=256, Lx=10, Ld=4)
    report = {}
    for bf16, tol in ((True, 2e-3), (False, 2e-3)):
        ref = O.render_rays(ro.cpu(), rd.cpu(), lambda x: O.flexible_forward(sdc, x, bf16=bf16),
                            lambda x: O.flexible_forward(sdf, x, bf16=bf16), opts, thr)
        errs = {}
        for name, a, b in zip(NAMES, res[:6], ref[:6]):
            errs[name] = float((a.reshape(b.shape).cpu() - b).abs().max())
        dex = torch.stack(res[6:], 0).reshape(20, -1).cpu()
        rdex = torch.stack(ref[6:], 0)
        spacing = (4.0 - 0.3) / 383.0
        errs["dex_same"] = float(((dex - rdex).abs() <= 1e-5).float().mean())
        errs["dex_within_one_spacing"] = float(((dex - rdex).abs() <= spacing).float().mean())
        report["bf16_oracle" if bf16 else "fp32_oracle"] = errs
        for name in ("rgb_c", "rgb_f", "acc_c", "acc_f"):
            assert errs[name] < tol, (bf16, name, errs)
        assert errs["depth_f"] < 2 * (4.0 - 0.3) / 127.0 and errs["dex_within_one_spacing"] > 0.97, (bf16, errs)
    print("\n[c5] tensor-core path:", json.dumps(report))
    _record("c5_128_256", report)
    assert float(res[5].mean()) > 0.05


def test_c2_64_128_on_the_tensor_core_path_2e3():
    """BASELINE config 2's networks and sampling on 256 rays of the 800x800 camera: rgb / acc within 2e-3 max abs of
    the fp32 oracle (the tolerance BASELINE.json states for bf16 operands), which is also what smoke() asserts."""
    torch.manual_seed(42)
    mc, mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4), nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    with torch.no_grad():
        for m in (mc, mf):
            m.fc_alpha.weight.mul_(150.0)
            m.fc_alpha.bias.fill_(2.0)
    sdc = {k: v.detach().clone() for k, v in mc.state_dict().items()}
    sdf = {k: v.detach().clone() for k, v in mf.state_dict().items()}
    T = O.pose_spherical_world2cam(30.0, -30.0, 4.0)
    K = torch.tensor([[1111.1, 0, 400.0], [0, 1111.1, 400.0], [0, 0, 1]])
    thr = [float(m) for m in range(5, 105, 5)]
    with torch.no_grad():
        ro, rd = nerf.get_ray_bundle(800, 800, None, T.cuda(), K.cuda(), row_start=400, row_count=1)
        ro, rd = ro[:, 272:528].contiguous(), rd[:, 272:528].contiguous()
        res = nerf.run_one_iter_of_nerf(800, 800, 1111.1, mc.cuda(), mf.cuda(), ro, rd, make_cfg(64, 128, 2.0, 6.0),
                                        mode="validation", encode_position_fn=nerf.get_embedding_function(10, True, True),
                                        encode_direction_fn=nerf.get_embedding_function(4, True, True), m_thres_cand=thr)
    opts = O.RenderOptions(near=2.0, far=6.0, num_coarse=64, num_fine=128, Lx=10, Ld=4)
    ref = O.render_rays(ro.cpu(), rd.cpu(), lambda x: O.flexible_forward(sdc, x), lambda x: O.flexible_forward(sdf, x), opts, thr)
    errs = {n: float((a.reshape(b.shape).cpu() - b).abs().max()) for n, a, b in zip(NAMES, res[:6], ref[:6])}
    print("\n[c2] tensor-core path vs fp32 oracle:", json.dumps(errs))
    _record("c2_64_128", errs)
    for name in ("rgb_c", "rgb_f", "acc_c", "acc_f"):
        assert errs[name] < 2e-3, (name, errs)
    assert errs["depth_f"] < (6.0 - 2.0) / 63.0
    assert float(res[5].mean()) > 0.05
