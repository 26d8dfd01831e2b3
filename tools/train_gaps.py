"""GPU idle time inside one C4 training iteration of nerf.Trainer: kernel timeline from torch.profiler (CUPTI), gaps
between consecutive kernels.  Run on the GPU box:  python tools/train_gaps.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import bench  # noqa: E402
import nerf  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    mc, mf = bench.state_dicts()
    mc, mf = mc.to(dev), mf.to(dev)
    mode = dict(chunksize=1 << 30, perturb=True, num_coarse=64, num_fine=128, white_background=False,
                radiance_field_noise_std=0.2, lindisp=False)
    cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=2.0, far=6.0),
                            nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    T, K = bench.camera()
    ro, rd = nerf.get_ray_bundle(bench.H, bench.W, None, T.to(dev), K.to(dev))
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    sel = torch.from_numpy(np.random.RandomState(0).choice(bench.H * bench.W, size=4096, replace=False)).to(dev)
    tgt = torch.rand(4096, 3, device=dev)
    trainer = nerf.Trainer(mc, mf, cfg, ex, ed, lr=5e-3)

    def step():
        return trainer.step(ro[sel], rd[sel], tgt)
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 3
    evs = evs[per:2 * per]                      # the middle iteration
    t0, prev_end, busy = evs[0].time_range.start, None, 0.0
    print("%9s %9s %8s  kernel" % ("start us", "dur us", "gap us"))
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        gap = (s - prev_end) if prev_end is not None else 0.0
        print("%9.1f %9.1f %8.1f  %s" % (s - t0, en - s, gap, e.name[:70]))
        busy += en - s
        prev_end = max(prev_end or en, en)
    span = prev_end - t0
    print("iteration: %d launches, span %.1f us, busy %.1f us, idle %.1f us (%.1f %%)" % (len(evs), span, busy, span - busy,
                                                                                          100 * (span - busy) / span))


if __name__ == "__main__":
    main()
