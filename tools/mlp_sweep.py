"""Tensor-core MLP kernel in isolation across the network shapes users of the reference actually run:
8x256 skip 4 (the paper configuration the YAMLs ask for), 8x128 skip 3 (messytable YAML) and 4x128 (what every
script instantiates, because num_layers / hidden_size are not forwarded - SURVEY.md section 8a-3).
CUDA-event time of one query over a frame of samples, TFLOP/s against the measured bf16 peaks.

    python tools/mlp_sweep.py [--rays 640000] [--samples 192]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf                                                     # noqa: E402
from nerf.train_utils import query_field                        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=640000)
    ap.add_argument("--samples", type=int, default=192)
    args = ap.parse_args()
    n, S = args.rays, args.samples
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    g = torch.Generator(device="cuda").manual_seed(0)
    ro = torch.randn(n, 3, device="cuda", generator=g)
    rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
    z = torch.sort(2 + 4 * torch.rand(n, S, device="cuda", generator=g), dim=-1).values
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    for name, kw in (("8x256 skip 4", dict(num_layers=8, hidden_size=256, skip_connect_every=4)),
                     ("8x128 skip 3", dict(num_layers=8, hidden_size=128, skip_connect_every=3)),
                     ("4x128 (as run)", dict()), ("PaperNeRFModel", None)):
        torch.manual_seed(0)
        m = (nerf.PaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4) if kw is None
             else nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4, **kw)).cuda()
        flop = 2 * sum(p.numel() for k, p in m.named_parameters() if k.endswith("weight"))
        with torch.no_grad():
            for _ in range(2):
                query_field(m, ro, rd, rd, z, ex, ed)
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); query_field(m, ro, rd, rd, z, ex, ed); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        tf = n * S * flop / ms / 1e9
        print("%-16s %8.2f ms  %7.1f M samples/s  %7.1f TFLOP/s  %.2f of sustained (%.0f)  %.2f of burst (%.0f)  "
              "%.0f GB/s of sample I/O" % (name, ms, n * S / ms / 1e3, tf, tf / peak["bf16_tflops_sustained"],
                                          peak["bf16_tflops_sustained"], tf / peak["bf16_tflops"], peak["bf16_tflops"],
                                          n * S * 20 / ms / 1e6))


if __name__ == "__main__":
    main()
