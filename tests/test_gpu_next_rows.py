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
    # exactly as the reference script calls it (train_dexnerf_rgb.py:391-404): gt_depth_torch = depth_target.cpu(),
    # pred_depth_torch = depth_fine_dex[cand].detach().cpu(), a CPU bool mask
    gt_cpu, mask_cpu = gt.cpu(), ((gt > 0) & (gt < 1.25)).cpu()
    for cand in (0, 3):
        e_cpu = nerf.compute_err_metric(gt_cpu, planes[cand].detach().cpu(), mask_cpu)
        np.testing.assert_allclose([e_cpu[k] for k in ("depth_abs_err", "depth_err2", "depth_err4", "depth_err8")],
                                   g["metric_errs"][cand], rtol=2e-5, atol=1e-7)


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


SCRIPT_YAML = """
experiment: {id: flow, logdir: logs, randomseed: 42, train_iters: 40, validate_every: 20, save_every: 20, print_every: 10}
dataset: {type: messytable, basedir: X, half_res: True, testskip: 1, no_ndc: True, near: 0.3, far: 4}
models:
  coarse: {type: FlexibleNeRFModel, num_layers: 8, hidden_size: 128, skip_connect_every: 3, include_input_xyz: True,
           log_sampling_xyz: True, num_encoding_fn_xyz: 10, use_viewdirs: True, include_input_dir: True,
           num_encoding_fn_dir: 4, log_sampling_dir: True}
  fine: {type: FlexibleNeRFModel, num_layers: 8, hidden_size: 128, skip_connect_every: 3, include_input_xyz: True,
         log_sampling_xyz: True, num_encoding_fn_xyz: 10, use_viewdirs: True, include_input_dir: True,
         num_encoding_fn_dir: 4, log_sampling_dir: True}
optimizer: {type: Adam, lr: 5.0E-3}
scheduler: {lr_decay: 250, lr_decay_factor: 0.1}
nerf:
  use_viewdirs: True
  encode_position_fn: positional_encoding
  encode_direction_fn: positional_encoding
  train: {num_random_rays: 256, chunksize: 16384, perturb: True, num_coarse: 64, num_fine: 64, white_background: False,
          radiance_field_noise_std: 0.2, lindisp: False}
  validation: {chunksize: 16384, perturb: False, num_coarse: 64, num_fine: 64, white_background: False,
               radiance_field_noise_std: 0., lindisp: False}
"""


def test_training_script_flow_on_a_messytable_dataset(tmp_path):
    """The statements of train_dexnerf_rgb.py, in its order and with its names, on a synthetic dataset in the
    messytable format: YAML -> CfgNode (:36-40), load_messytable_data (:61-67), encoders and models built by
    name with the five keyword arguments the script forwards (:106-140 - hence 4x128 networks whatever the
    YAML says), torch.optim.Adam (:142-148), per iteration get_ray_bundle / meshgrid_xy / np.random.choice ray
    selection / run_one_iter_of_nerf(mode="train") / two mse losses / backward / step / exponential learning
    rate (:221-289), then the validation block (:317-428): full-image render under no_grad, the Dex threshold
    selection against the ground-truth depth, the error colour image, and the checkpoint dict (:442-457)."""
    import yaml
    import dataset_fixture as DF
    from nerf import (CfgNode, compute_err_metric, depth_error_img, get_embedding_function, get_ray_bundle, img2mse,
                      load_messytable_data, meshgrid_xy, models, mse2psnr, run_one_iter_of_nerf)
    cfg = CfgNode(yaml.safe_load(SCRIPT_YAML))
    cfg.dataset.basedir = DF.build_messytable(str(tmp_path / "messy"))
    images, poses, render_poses, hwf, i_split, intrinsics, depths = load_messytable_data(
        cfg.dataset.basedir, half_res=cfg.dataset.half_res, testskip=cfg.dataset.testskip)
    i_train, i_val, i_test = i_split
    H, W, focal = int(hwf[0]), int(hwf[1]), hwf[2]
    # the fixture's intrinsics describe a 1920x1080 sensor; point them at the 32x18 images it really holds
    intrinsics[:, 0, 0] = intrinsics[:, 1, 1] = 40.0
    intrinsics[:, 0, 2], intrinsics[:, 1, 2] = W / 2, H / 2
    np.random.seed(cfg.experiment.randomseed)
    torch.manual_seed(cfg.experiment.randomseed)
    device = "cuda"
    encode_position_fn = get_embedding_function(num_encoding_functions=cfg.models.coarse.num_encoding_fn_xyz,
                                                include_input=cfg.models.coarse.include_input_xyz,
                                                log_sampling=cfg.models.coarse.log_sampling_xyz)
    encode_direction_fn = get_embedding_function(num_encoding_functions=cfg.models.coarse.num_encoding_fn_dir,
                                                 include_input=cfg.models.coarse.include_input_dir,
                                                 log_sampling=cfg.models.coarse.log_sampling_dir)
    nets = []
    for which in ("coarse", "fine"):
        assert hasattr(cfg.models, which)
        c = getattr(cfg.models, which)
        nets.append(getattr(models, c.type)(num_encoding_fn_xyz=c.num_encoding_fn_xyz, num_encoding_fn_dir=c.num_encoding_fn_dir,
                                            include_input_xyz=c.include_input_xyz, include_input_dir=c.include_input_dir,
                                            use_viewdirs=c.use_viewdirs).to(device))
    model_coarse, model_fine = nets
    assert model_coarse.layer1.weight.shape == (128, 63) and len(model_coarse.layers_xyz) == 3    # the as-run 4x128
    optimizer = getattr(torch.optim, cfg.optimizer.type)(list(model_coarse.parameters()) + list(model_fine.parameters()),
                                                         lr=cfg.optimizer.lr)
    m_thres_cand = np.arange(5, 105, 5)
    # targets a NeRF can fit: the training view rendered by a frozen teacher pair (the fixture's pixels are noise)
    torch.manual_seed(7)
    teacher = [models.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4).to(device) for _ in range(2)]
    with torch.no_grad():
        for tm in teacher:
            tm.fc_alpha.weight.mul_(300.0)
    losses = []
    for i in range(cfg.experiment.train_iters):
        model_coarse.train(); model_fine.train()
        img_idx = np.random.choice(i_train)
        pose_target = poses[img_idx, :, :].to(device)
        intrinsic_target = intrinsics[img_idx, :, :].to(device)
        ray_origins, ray_directions = get_ray_bundle(H, W, focal, pose_target, intrinsic_target)
        if i == 0:
            with torch.no_grad():
                img_target = run_one_iter_of_nerf(H, W, intrinsic_target[0, 0], teacher[0], teacher[1], ray_origins,
                                                  ray_directions, cfg, mode="validation",
                                                  encode_position_fn=encode_position_fn,
                                                  encode_direction_fn=encode_direction_fn, m_thres_cand=m_thres_cand)[3]
            assert img_target.shape == (H, W, 3)
        coords = torch.stack(meshgrid_xy(torch.arange(H).to(device), torch.arange(W).to(device)), dim=-1).reshape((-1, 2))
        select_inds = np.random.choice(coords.shape[0], size=(cfg.nerf.train.num_random_rays), replace=False)
        select_inds = coords[select_inds]
        ro = ray_origins[select_inds[:, 0], select_inds[:, 1], :]
        rd = ray_directions[select_inds[:, 0], select_inds[:, 1], :]
        target_s = img_target[select_inds[:, 0], select_inds[:, 1]]
        nerf_out = run_one_iter_of_nerf(H, W, intrinsic_target[0, 0], model_coarse, model_fine, ro, rd, cfg, mode="train",
                                        encode_position_fn=encode_position_fn, encode_direction_fn=encode_direction_fn,
                                        m_thres_cand=m_thres_cand)
        assert len(nerf_out) == 6 + len(m_thres_cand)
        rgb_coarse, rgb_fine = nerf_out[0], nerf_out[3]
        loss = torch.nn.functional.mse_loss(rgb_coarse[..., :3], target_s[..., :3]) + \
            torch.nn.functional.mse_loss(rgb_fine[..., :3], target_s[..., :3])
        loss.backward()
        psnr = mse2psnr(loss.item())
        optimizer.step()
        optimizer.zero_grad()
        lr_new = cfg.optimizer.lr * (cfg.scheduler.lr_decay_factor ** (i / (cfg.scheduler.lr_decay * 1000)))
        for param_group in optimizer.param_groups:
            param_group["lr"] = lr_new
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.isfinite(psnr)
    assert np.mean(losses[-5:]) < 0.5 * np.mean(losses[:5]), (losses[:5], losses[-5:])

    # ---- validation block
    model_coarse.eval(); model_fine.eval()
    with torch.no_grad():
        img_idx = np.random.choice(i_val)
        pose_target = poses[img_idx, :, :].to(device)
        intrinsic_target = intrinsics[img_idx, :, :].to(device)
        depth_target = depths[img_idx].to(device)
        ray_origins, ray_directions = get_ray_bundle(H, W, focal, pose_target, intrinsic_target)
        out = run_one_iter_of_nerf(H, W, intrinsic_target[0, 0], model_coarse, model_fine, ray_origins, ray_directions, cfg,
                                   mode="validation", encode_position_fn=encode_position_fn,
                                   encode_direction_fn=encode_direction_fn, m_thres_cand=m_thres_cand)
        rgb_fine, depth_fine_dex = out[3], list(out[6:])
        assert rgb_fine.shape == (H, W, 3) and all(d.shape == (H, W) for d in depth_fine_dex)
        val_loss = img2mse(rgb_fine[..., :3], images[img_idx].to(device)[..., :3])
        assert np.isfinite(mse2psnr(val_loss.item()))
        img_ground_mask = (depth_target > 0) & (depth_target < 1.25)
        min_err, min_abs_err, min_abs_depth, min_cand = None, 1000.0, None, 0
        for cand in range(m_thres_cand.shape[0]):
            err = compute_err_metric(depth_target, depth_fine_dex[cand], img_ground_mask)
            if err["depth_abs_err"] < min_abs_err:
                min_abs_err, min_err, min_abs_depth, min_cand = err["depth_abs_err"], err, depth_fine_dex[cand], m_thres_cand[cand]
        assert min_err is not None and set(min_err) == {"depth_abs_err", "depth_err2", "depth_err4", "depth_err8"}
        # the one-launch selection of this repo finds the same candidate
        idx, err2, _ = nerf.select_dex_threshold(torch.stack(depth_fine_dex).reshape(len(depth_fine_dex), -1),
                                                 depth_target.reshape(-1))
        assert m_thres_cand[idx] == min_cand and abs(err2["depth_abs_err"] - min_abs_err) < 1e-3
        err_img = depth_error_img(min_abs_depth.unsqueeze(0) * 1000, depth_target.unsqueeze(0) * 1000,
                                  img_ground_mask.unsqueeze(0))
        assert err_img.shape == (H, W, 3)
    checkpoint_dict = {"iter": i, "model_coarse_state_dict": model_coarse.state_dict(),
                       "model_fine_state_dict": model_fine.state_dict(), "optimizer_state_dict": optimizer.state_dict(),
                       "loss": loss, "psnr": psnr}
    torch.save(checkpoint_dict, str(tmp_path / "checkpoint.ckpt"))
    back = torch.load(str(tmp_path / "checkpoint.ckpt"), weights_only=False)
    model_fine.load_state_dict(back["model_fine_state_dict"])
    assert isinstance(cfg.dump(), str)
