"""CUDA-event timing of the HBM-side kernels of a render step in isolation (compositing, resample +
merge) at the BASELINE sizes, against their algorithmic bytes and the
measured HBM peak.  Inputs are larger than the 126 MB L2, so every launch streams from HBM.

    python tools/small_kernels.py [--rays 640000]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf                                                     # noqa: E402,F401
from nerf import _lib as L                                      # noqa: E402
from nerf.volume_rendering_utils import render_maps             # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=640000)
    args = ap.parse_args()
    n = args.rays
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = float(peak["hbm_gbs"])
    g = torch.Generator(device="cuda").manual_seed(0)
    thr = torch.linspace(5, 100, 20, device="cuda")
    rd = torch.randn(n, 3, device="cuda", generator=g)
    rows = []
    for S in (64, 192, 128, 384):
        rf = torch.randn(n, S, 4, device="cuda", generator=g)
        # section 8d value distribution: sigma_raw ~ 30 N(0,1) with 5 % spikes U(50,300) ("peaky"), or the
        # near-zero field of a random-init network ("flat": what the C2 bench step sees)
        z = torch.sort(torch.rand(n, S, device="cuda", generator=g) * 4 + 2, dim=-1).values
        for name in ("flat", "peaky"):
            if name == "flat":
                rf[..., 3] = 0.05 * torch.randn(n, S, device="cuda", generator=g)
            else:
                sig = 30 * torch.randn(n, S, device="cuda", generator=g)
                spike = torch.rand(n, S, device="cuda", generator=g) < 0.05
                sig[spike] = 50 + 250 * torch.rand(int(spike.sum()), device="cuda", generator=g)
                rf[..., 3] = sig
            for want_w in (True, False):
                nbytes = n * (S * (20 + (4 if want_w else 0)) + 36 + 4 * 20)
                ms = timed(lambda: render_maps(rf, z, rd, None, False, thr, 20, want_weights=want_w))
                rows.append(("composite S=%d %s weights=%d" % (S, name, want_w), ms, nbytes))
        del rf
    for Nc, Nf in ((64, 128), (128, 256)):
        zc = torch.sort(torch.rand(n, Nc, device="cuda", generator=g) * 4 + 2, dim=-1).values
        for name in ("flat", "peaky"):
            w = torch.rand(n, Nc, device="cuda", generator=g)
            w = w * 0.02 if name == "flat" else w ** 8
            for uname in ("det", "rand"):
                u = None if uname == "det" else torch.rand(n, Nf, device="cuda", generator=g)
                zf = torch.empty(n, Nc + Nf, device="cuda")
                nbytes = n * (8 * Nc + 4 * (Nc + Nf) + (4 * Nf if u is not None else 0))

                def run():
                    L.check(L.lib().dexnerf_resample_merge(L.ptr(zc), L.ptr(w), n, Nc, Nf, L.ptr(u), L.ptr(zf),
                                                           L.stream_ptr()), "resample_merge")
                ms = timed(run)
                rows.append(("resample_merge %d+%d %s u=%s" % (Nc, Nf, name, uname), ms, nbytes))
    for name, ms, nbytes in rows:
        gbs = nbytes / ms / 1e6
        print("%-44s %8.3f ms  %8.1f MB  %7.0f GB/s  %.2f of HBM peak (%.0f)" % (name, ms, nbytes / 1e6, gbs, gbs / hbm, hbm))


if __name__ == "__main__":
    main()
