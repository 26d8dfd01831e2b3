"""GPU idle time inside one C2 render step: kernel timeline from torch.profiler (CUPTI), gaps between
consecutive kernels.  Run on the GPU box:  python tools/step_gaps.py"""
import os
import sys

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
    cfg = bench.make_cfg(nerf)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    T, K = bench.camera()
    T, K = T.to(dev), K.to(dev)

    def step():
        ro, rd = nerf.get_ray_bundle(bench.H, bench.W, None, T, K)
        with torch.no_grad():
            return nerf.run_one_iter_of_nerf(bench.H, bench.W, bench.FX, mc, mf, ro, rd, cfg, mode="validation",
                                             encode_position_fn=ex, encode_direction_fn=ed,
                                             m_thres_cand=bench.THRESHOLDS)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    prev_end = None
    busy = 0.0
    rows = []
    for e in evs:
        s, en = e.time_range.start, e.time_range.end
        gap = (s - prev_end) if prev_end is not None else 0.0
        rows.append((s - t0, en - s, gap, e.name[:60]))
        busy += en - s
        prev_end = max(prev_end or en, en)
    span = prev_end - t0
    print("span %.3f ms, busy %.3f ms, idle %.3f ms over %d device activities" % (span / 1e3, busy / 1e3, (span - busy) / 1e3, len(evs)))
    for r in rows:
        if r[2] > 30 or r[1] > 300:
            print("t=%9.1f us  dur=%9.1f us  gap_before=%8.1f us  %s" % r)


if __name__ == "__main__":
    main()
