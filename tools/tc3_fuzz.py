"""Randomised agreement of the three-tile hidden-128 kernel (csrc/mlp_tc3.cu) with the pair kernel (the debug-tap form
of dexnerf_tc_query) over network shapes and sample counts, each case run twice (bit-identical or a hand-off raced).
    python tools/tc3_fuzz.py [cases] [seed]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
from nerf import tensorcore

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = torch.Generator().manual_seed(seed)
ri = lambda lo, hi: int(torch.randint(lo, hi + 1, (1,), generator=g))
worst_rgb = worst_sig = 0.0
t0 = time.time()
for c in range(cases):
    layers, skip, Lx = ri(2, 12), ri(1, 6), (10, 6, 4)[ri(0, 2)]
    viewdirs = ri(0, 3) > 0
    S = (7, 33, 64, 128, 192, 100)[ri(0, 5)]
    n = (1, 3, 50, 777, 2500, 148 * 6 + 1)[ri(0, 5)]
    torch.manual_seed(seed * 1000 + c)
    model = nerf.FlexibleNeRFModel(layers, 128, skip, Lx, 4, use_viewdirs=viewdirs).cuda()
    ex, ed = nerf.get_embedding_function(Lx, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    ro = (torch.randn(n, 3, generator=g) * 0.3).cuda()
    rd = torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    a, b, w = (torch.full((n, S, 4), float("nan"), device="cuda") for _ in range(3))
    tensorcore.query(model, prog, ro, rd, vd, z, a)
    tensorcore.query(model, prog, ro, rd, vd, z, b)
    dbg = torch.zeros(n * S * 128, device="cuda")
    tensorcore.query(model, prog, ro, rd, vd, z, w, dbg=dbg, dbg_layer=0, dbg_pass=0)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all(), (c, "nan")
    assert torch.equal(a, b), (c, layers, skip, Lx, viewdirs, n, S, "two runs differ")
    e_rgb = float((a[..., :3] - w[..., :3]).abs().max()) / max(1.0, float(w[..., :3].abs().max()))
    e_sig = float((a[..., 3] - w[..., 3]).abs().max()) / max(1.0, float(w[..., 3].abs().max()))
    assert e_rgb < 1e-3 and e_sig < 4e-3, (c, layers, skip, Lx, viewdirs, n, S, e_rgb, e_sig)
    worst_rgb, worst_sig = max(worst_rgb, e_rgb), max(worst_sig, e_sig)
print("%d cases ok in %.1f s; worst rgb diff %.2e, worst sigma diff %.2e (relative to max(1, largest value))" % (cases, time.time() - t0, worst_rgb, worst_sig))
# a long launch (many groups per CTA) five times: bit-identical
torch.manual_seed(5)
model = nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda()
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
prog = model.program(ex, ed)
n, S = 40000, 192
ro = (torch.randn(n, 3, generator=g) * 0.3).cuda(); rd = torch.randn(n, 3, generator=g).cuda()
vd = rd / rd.norm(dim=-1, keepdim=True)
z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
outs = []
for _ in range(5):
    o = torch.empty(n, S, 4, device="cuda")
    tensorcore.query(model, prog, ro, rd, vd, z, o)
    outs.append(o)
torch.cuda.synchronize()
assert all(torch.equal(outs[0], o) for o in outs[1:]), "long launch: runs differ"
print("long launch (60 000 tiles) five times: bit-identical")
