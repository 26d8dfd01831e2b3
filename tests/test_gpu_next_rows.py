"""The callers on either side of the hot path: the evaluation render loop
over a camera path and the validation-time Dex-depth error metrics, on the GPU."""
import numpy as np
import pytest
import torch

import nerf
from oracle import nerf_oracle as O

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def test_depth_error_metrics_golden(golden):
    """compute_err_metric / the threshold selection loop against the reference's own outputs."""
    g = golden("next_rows")
    gt, planes, mask = t(g["metric_gt"]).cuda(), t(g["metric_planes"]).cuda(), t(g["metric_mask"]).cuda()
    table, best = nerf.dex_depth_error_metrics(planes, gt)            # reference mask rule built in
    np.testing.assert_allclose(table.cpu().numpy(), g["metric_errs"], rtol=2e-5, atol=1e-7)
    assert int(best) == int(g["metric_best"])
    table2, _ = nerf.dex_depth_error_metrics(list(planes), gt, mask)   # explicit mask, sequence of planes
    assert torch.equal(table, table2)
    e = nerf.compute_err_metric(gt, planes[3], mask)
    np.testing.assert_allclose([e["depth_abs_err"], e["depth_err2"], e["depth_err4"], e["depth_err8"]],
                               g["metric_errs"][3], rtol=2e-5, atol=1e-7)
    idx, err, tab = nerf.select_dex_threshold(planes, gt)
    assert idx == int(g["metric_best"]) and abs(err["depth_abs_err"] - g["metric_errs"][idx, 0]) < 1e-4


def test_depth_error_metrics_full_frame_vs_oracle():
    """1280x720, T=20 (BASELINE config 5 size) against the CPU oracle; nothing crosses -> index -1."""
    g = torch.Generator().manual_seed(3)
    H, W, T = 720, 1280, 20
    gt = 0.2 + 1.3 * torch.rand(H, W, generator=g)
    planes = gt[None] + 0.01 * torch.randn(T, H, W, generator=g)
    planes[11] = gt + 0.001 * torch.randn(H, W, generator=g)
    idx, err, table = nerf.select_dex_threshold(planes.cuda(), gt.cuda())
    obest, oerr = O.select_dex_threshold(list(planes), gt)
    assert idx == obest == 11
    for k in ("depth_abs_err", "depth_err2", "depth_err4", "depth_err8"):
        assert abs(err[k] - oerr[k]) <= 2e-5 * max(1.0, abs(oerr[k]))
    far = planes + 5.0                                # every candidate is off by metres: abs err > 1000 mm
    idx, err, _ = nerf.select_dex_threshold(far.cuda(), gt.cuda())
    assert idx == -1 and err is None


def test_pose_spherical_and_render_path(golden):
    g = golden("next_rows")
    for args, ref in zip(g["pose_args"], g["poses"]):
        np.testing.assert_allclose(nerf.pose_spherical(*map(float, args)).numpy(), ref, rtol=1e-6, atol=1e-6)
    # the blender -> fork conversion agrees with the oracle's helper used by the benchmarks
    w2c = nerf.world2cam_from_blender_pose(nerf.pose_spherical(30.0, -30.0, 4.0))
    np.testing.assert_allclose(w2c.numpy(), O.pose_spherical_world2cam(30.0, -30.0, 4.0).numpy(), rtol=1e-5, atol=1e-6)
    assert nerf.render_poses_spherical(8).shape == (8, 4, 4)

    torch.manual_seed(0)
    mc, mf = nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda(), nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda()
    mode = dict(chunksize=4096, perturb=False, num_coarse=32, num_fine=32, white_background=True,
                radiance_field_noise_std=0.0, lindisp=False)
    cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=2.0, far=6.0),
                            nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    H, W = 20, 28
    K = torch.tensor([[30.0, 0, 14.0], [0, 30.0, 10.0], [0, 0, 1]])
    poses = [nerf.world2cam_from_blender_pose(p) for p in nerf.render_poses_spherical(3)]
    thr = [5.0, 50.0]
    seen = []
    frames, secs = nerf.render_path(poses, H, W, K, mc, mf, cfg, ex, ed, m_thres_cand=thr,
                                    on_frame=lambda i, out: seen.append(i))
    assert seen == [0, 1, 2] and len(secs) == 3 and all(s > 0 for s in secs)
    for pose, out in zip(poses, frames):
        with torch.no_grad():
            ro, rd = nerf.get_ray_bundle(H, W, None, pose.cuda(), K.cuda())
            ref = nerf.run_one_iter_of_nerf(H, W, 30.0, mc, mf, ro, rd, cfg, mode="validation", encode_position_fn=ex,
                                            encode_direction_fn=ed, m_thres_cand=thr)
        assert len(out) == 8 and out[3].shape == (H, W, 3)
        for a, b in zip(out, ref):
            assert torch.equal(a, b)
    # the ranks of a 4-GPU render, one after the other: their row blocks tile the frame
    parts = [nerf.render_path(poses[:1], H, W, K, mc, mf, cfg, ex, ed, m_thres_cand=thr, rank=r, world=4)[0][0]
             for r in range(4)]
    for k in range(8):
        assert torch.equal(torch.cat([p[k] for p in parts], 0), frames[0][k])
    img = nerf.cast_to_image(frames[0][3])
    assert img.shape == (H, W, 3) and img.dtype == np.uint8
