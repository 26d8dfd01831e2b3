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


def test_ray_cache_files_round_trip(tmp_path):
    """The reference's ray-cache format (cache_dataset.py:104-135) written from this repo's ray generator and
    consumed the way train_dexnerf_rgb.py:186-204 does."""
    from nerf import cache_utils as CU
    H, W = 12, 16
    K = torch.tensor([[20.0, 0, 8.0], [0, 20.0, 6.0], [0, 0, 1]])
    pose = nerf.world2cam_from_blender_pose(nerf.pose_spherical(40.0, -25.0, 4.0))
    image = torch.rand(H, W, 3)
    full = CU.train_cache_entry(H, W, 20.0, pose, K, image)                        # --sample-all
    assert set(full) == {"height", "width", "focal_length", "ray_bundle", "target"}
    assert full["ray_bundle"].shape == (2, H, W, 3) and not full["ray_bundle"].is_cuda
    ro, rd = nerf.get_ray_bundle(H, W, None, pose.cuda(), K.cuda())
    assert torch.equal(full["ray_bundle"][0], ro.cpu()) and torch.equal(full["ray_bundle"][1], rd.cpu())
    sub = CU.train_cache_entry(H, W, 20.0, pose, K, image, num_random_rays=50, rng=np.random.RandomState(3))
    assert sub["ray_bundle"].shape == (2, 50, 3) and sub["target"].shape == (50, 3)
    val = CU.val_cache_entry(H, W, 20.0, pose, K, image)
    assert set(val) == {"height", "width", "focal_length", "ray_origins", "ray_directions", "target"}
    path = tmp_path / "0000.data"
    CU.save_cache_entry(full, path)
    back = CU.load_cache_entry(path)
    o, d, tg = CU.training_rays_from_cache(back, 32, rng=np.random.RandomState(1))
    assert o.shape == (32, 3) and o.is_cuda and tg.shape == (32, 3)
    sel = torch.from_numpy(np.random.RandomState(1).choice(H * W, size=(32,), replace=False))
    assert torch.equal(d.cpu(), rd.reshape(-1, 3).cpu()[sel]) and torch.equal(tg.cpu(), image.reshape(-1, 3)[sel])
