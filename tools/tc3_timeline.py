"""Timeline of the three-tile hidden-128 kernel (csrc/mlp_tc3.cu): clock stamps of the issuer, the scout and the
epilogue warps of CTA 0 over a window of passes in steady state.  Needs the bring-up build:
    python dex-nerf_b200/build.py --define DEXNERF_TC3_BRINGUP=1 --out libdexnerf_bringup.so
    DEXNERF_LIB=$PWD/dex-nerf_b200/lib/libdexnerf_bringup.so python tools/tc3_timeline.py [8x128|4x128] [group]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
from nerf import tensorcore

shape = sys.argv[1] if len(sys.argv) > 1 else "8x128"
group = int(sys.argv[2]) if len(sys.argv) > 2 else 6
model = (nerf.FlexibleNeRFModel(8, 128, 3, 10, 4) if shape == "8x128"
         else nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)).cuda()
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
prog = model.program(ex, ed)
nl = len(model.layers_xyz) + 3
n, S = 148 * 3 * 16, 128                     # 16 groups per CTA
g = torch.Generator().manual_seed(1)
ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
vd = rd / rd.norm(dim=-1, keepdim=True)
z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
rf = torch.empty(n, S, 4, device="cuda")
tensorcore.query(model, prog, ro, rd, vd, z, rf)
torch.cuda.synchronize()
buf = torch.zeros(32 * 512, dtype=torch.int32, device="cuda")
lo = group * nl * 3
os.environ["DEXNERF_TC3_TRACE"] = str(buf.data_ptr())
os.environ["DEXNERF_TC3_WINDOW"] = "%d,%d" % (lo, lo + nl * 3)
tensorcore.query(model, prog, ro, rd, vd, z, rf)
torch.cuda.synchronize()
tr = buf.cpu().view(32, 512)
ev = {}
for role in range(32):
    cnt = int(tr[role, 0])
    for i in range(cnt):
        w, c = int(tr[role, 1 + 2 * i]) & 0xFFFFFFFF, int(tr[role, 2 + 2 * i]) & 0xFFFFFFFF
        ev.setdefault((w & 0xFFFF, role), {})[w >> 16] = c
t0 = min(c for d in ev.values() for c in d.values())
rel = lambda c: (c - t0) & 0xFFFFFFFF
print("%s: %d layers, passes %d..%d of CTA 0 (group %d); cycles relative to the first stamp" % (shape, nl, lo, lo + nl * 3, group))
print("producer (per layer: chunk c -> [slot wait begins, slot free = load issued]):")
for l in range(nl):
    d = ev.get((l * 3, 0), {})
    print("  layer %2d: " % l + "  ".join("c%d %s -> %s" % (c, ("%6d" % rel(d[2 * c])) if 2 * c in d else "-", ("%6d" % rel(d[2 * c + 1])) if 2 * c + 1 in d else "-") for c in range(3) if 2 * c in d))
print("pass layer tile acc | issuer of the tile: at the pass -> weights / enc ok -> its turn -> gate open -> issued | epilogue team: first warp sees D .. last warp done (last arithmetic end)")
prev_issue = None
for k in range(nl * 3):
    iss = ev.get((k, 1 + k % 3), {})
    team = (lo + k) & 1
    ws = [ev.get((k, 4 + team * 8 + w), {}) for w in range(8)]
    seen = [rel(d[1]) for d in ws if 1 in d]; done = [rel(d[3]) for d in ws if 3 in d]; arith = [rel(d[2]) for d in ws if 2 in d]
    f = lambda d, c: ("%6d" % rel(d[c])) if c in d else "     -"
    gap = "" if prev_issue is None or 1 not in iss else " (+%d since the previous pass was issued)" % (rel(iss[1]) - prev_issue)
    prev_issue = rel(iss[1]) if 1 in iss else prev_issue
    print("%3d  %3d  %3d  D%d | %s %s %s | %s %s %s | %s .. %s (%s)%s" % (
        k, k // 3, k % 3, team, f(iss, 0), f(iss, 4), f(iss, 2), f(iss, 3), "", f(iss, 1),
        ("%6d" % min(seen)) if seen else "     -", ("%6d" % max(done)) if done else "     -",
        ("%6d" % max(arith)) if arith else "-", gap))
