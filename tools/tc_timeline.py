"""Pipeline timeline of the tcgen05 MLP kernel (block 0, first tile pair): SM-clock stamps of each
MMA pass issue window and each epilogue pass, dumped through the kernel's timing tap
(dbg_layer = -2).  Run on the GPU box:  python tools/tc_timeline.py [n_rays] [S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))

import nerf  # noqa: E402
from nerf import tensorcore  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 16
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    tap = int(sys.argv[3]) if len(sys.argv) > 3 else 0        # which tile pair of CTA 0 is recorded (0 = first)
    torch.manual_seed(1)
    shape = sys.argv[4] if len(sys.argv) > 4 else "8x256"   # "8x256" | "8x128" | "4x128" (the as-run networks)
    model = {"8x256": lambda: nerf.FlexibleNeRFModel(8, 256, 4, 10, 4),
             "8x128": lambda: nerf.FlexibleNeRFModel(8, 128, 3, 10, 4),
             "4x128": lambda: nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)}[shape]().cuda()
    flop = 2 * sum(p.numel() for k, p in model.named_parameters() if k.endswith("weight"))
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    g = torch.Generator().manual_seed(0)
    ro = (torch.randn(n, 3, generator=g) * 0.3).cuda()
    rd = torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    rf = torch.empty(n, S, 4, device="cuda")
    for _ in range(2):
        tensorcore.query(model, prog, ro, rd, vd, z, rf)
    tl = torch.zeros(2048, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tensorcore.query(model, prog, ro, rd, vd, z, rf, dbg=tl.view(torch.float32), dbg_layer=-2, dbg_pass=tap)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    tiles = (n * S + 127) // 128
    print("kernel %.3f ms for %d samples (%d tiles, %.1f pairs/SM): %.2f ns/sample, %.1f TFLOP/s"
          % (ms, n * S, tiles, tiles / 2 / 148, ms * 1e6 / (n * S), n * S * flop / ms / 1e9))
    for mode, name in ((-3, "weights traffic cut to 1 KB/chunk (results invalid)"), (-4, "debug build, full traffic")):
        e0.record()
        tensorcore.query(model, prog, ro, rd, vd, z, rf, dbg=tl.view(torch.float32), dbg_layer=mode)
        e1.record()
        torch.cuda.synchronize()
        print("%s: %.3f ms" % (name, e0.elapsed_time(e1)))
    t = tl.cpu().tolist()
    t0 = min(x for x in t[:512] if x > 0)
    print("layer pass tile |  mma_start  mma_end (issue) |  epi_start  epi_end | epi_len  gap(epi_start - mma_end) | last warp's epi_end")
    for l in range(10):
        for p in range(2):
            for tile in range(2):
                i = ((l * 2 + p) * 2 + tile) * 2
                ms_, me_, es_, ee_ = t[i], t[i + 1], t[256 + i], t[256 + i + 1]
                if ms_ == 0:
                    continue
                last = t[800 + (l * 2 + p) * 2 + tile]
                print("%5d %4d %4d | %10d %9d | %10d %8d | %7d %6d | %8d"
                      % (l, p, tile, ms_ - t0, me_ - t0, es_ - t0, ee_ - t0, ee_ - es_, es_ - me_, last - t0 if last else -1))
    print("issuer: layer pass tile | before gate -> after gate -> first MMA issue (relative to the previous pass's issue end)")
    prev_end = None
    for l in range(10):
        for p in range(2):
            for tile in range(2):
                i = ((l * 2 + p) * 2 + tile)
                g0, g1 = t[900 + i * 3], t[900 + i * 3 + 1]
                ms_, me_ = t[i * 2], t[i * 2 + 1]
                if ms_ == 0:
                    continue
                if prev_end is not None and l in (2, 3, 4):
                    print("  %d %d %d | loop overhead %5d | gate wait %5d | weights wait + fence %5d | issue %5d"
                          % (l, p, tile, g0 - prev_end, g1 - g0, ms_ - g1, me_ - ms_))
                prev_end = me_
    for p in range(2):
        for tile in range(2):
            st = t[600 + (p * 2 + tile) * 16: 600 + (p * 2 + tile) * 16 + 11]
            if st[0]:
                print("layer 2 pass %d tile %d chunk stamps (rel):" % (p, tile), [x - st[0] if x else None for x in st])
    last = max(t[256:512])
    print("pair span: %d cycles (MMA floor for 2 tiles: 37120)" % (last - t0))


if __name__ == "__main__":
    main()
