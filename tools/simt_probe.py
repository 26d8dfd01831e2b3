"""Throughput of the fp32 CUDA-core MLP path (nerf.set_precision("fp32"), mlp_simt.cu) on 40 000 rays x 192
samples: the full-precision fallback of FlexibleNeRFModel and the only path of PaperNeRFModel.

    python tools/simt_probe.py
"""
import sys, os, torch
sys.path.insert(0, "/root/repo/dex-nerf_b200")
import nerf
from nerf.train_utils import query_field
nerf.set_precision("fp32")
n, S = 40000, 192
g = torch.Generator(device="cuda").manual_seed(0)
ro = torch.randn(n, 3, device="cuda", generator=g); rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(2 + 4 * torch.rand(n, S, device="cuda", generator=g), dim=-1).values
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
for name, ctor in (("flex 8x256", lambda: nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)), ("flex 4x128", lambda: nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)),
                   ("paper", lambda: nerf.PaperNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4))):
    m = ctor().cuda()
    flop = 2 * sum(p.numel() for k, p in m.named_parameters() if k.endswith("weight"))
    with torch.no_grad():
        query_field(m, ro, rd, rd, z, ex, ed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); query_field(m, ro, rd, rd, z, ex, ed); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%-12s fp32 SIMT: %.2f ms for %d samples, %.1f TFLOP/s" % (name, ms, n * S, n * S * flop / ms / 1e9))
