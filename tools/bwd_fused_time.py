"""Timing of the fused backward launch in isolation at BASELINE config 4's fine-pass size (4096 rays x 192 samples),
with bring-up variants: bit 3 (8) no discard, bit 4 (16) consumers do not wait for the flags (results invalid),
variant >> 8 = number of chain CTAs.   python tools/bwd_fused_time.py [rays] [S] [variant ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
from nerf import tensorcore, training

n, S = int(sys.argv[1]), int(sys.argv[2])
variants = [int(v) for v in sys.argv[3:]] or [0]
hidden = int(os.environ.get("HIDDEN", "256"))
torch.manual_seed(0)
model = (nerf.FlexibleNeRFModel(8, 256, 4, 10, 4) if hidden == 256 else nerf.FlexibleNeRFModel(8, 128, 3, 10, 4)).cuda()
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
prog = model.program(ex, ed); spec = tensorcore.spec_for(model, prog)
g = torch.Generator().manual_seed(1)
ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
vd = rd / rd.norm(dim=-1, keepdim=True)
z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
d_rf = (torch.randn(n, S, 4, generator=g) * 0.05).cuda()
rf, tape = training.query_train(model, prog, spec, ro, rd, vd, z)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(what, variant):
    ts = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=what, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


print("split dx: %.3f ms   dw: %.3f ms" % (timed(1, 0), timed(2, 0)))
for v in [int(x) for x in os.environ.get("DW_VARIANTS", "").split(",") if x]:
    print("stand-alone dW variant %d (CTAs %d, flags 0x%x): %.3f ms" % (v, v >> 8, v & 0xFF, timed(2, v)))
for v in variants:
    print("fused variant %d (chain %d, flags 0x%x): %.3f ms" % (v, v >> 8, v & 0xFF, timed(4, v)))
for v in [int(x) for x in os.environ.get("SHARED_VARIANTS", "0").split(",") if x]:
    print("shared-SM kernel variant 0x%x: %.3f ms" % (v, timed(8, v)))
want = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=3)
got = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=8)
torch.cuda.synchronize()
print("shared-SM kernel vs split: rel err %.3e" % float((got - want).norm() / want.norm()))
