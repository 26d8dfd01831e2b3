"""Small query on the three-tile hidden-128 kernel against the pair kernel (DEXNERF_TC3=0 in a second process is the
usual A/B; here: one process, results against the bf16 emulation of the oracle).  python tools/tc3_check.py [n] [S]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
from nerf import tensorcore
from oracle import nerf_oracle as O

TRACE = None
if os.environ.get("TC3_TRACE"):
    TRACE = torch.zeros(32 * 512, dtype=torch.int32).pin_memory()
    os.environ["DEXNERF_TC3_TRACE"] = str(TRACE.data_ptr())
n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 50, int(sys.argv[2]) if len(sys.argv) > 2 else 128
torch.manual_seed(0)
for name, model in (("8x128", nerf.FlexibleNeRFModel(8, 128, 3, 10, 4)), ("4x128", nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)),
                    ("8x128 nodirs", nerf.FlexibleNeRFModel(8, 128, 3, 10, 4, use_viewdirs=False))):
    model = model.cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    g = torch.Generator().manual_seed(1)
    ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    rf = torch.full((n, S, 4), float("nan"), device="cuda")
    tensorcore.query(model, prog, ro, rd, vd, z, rf)
    try:
        torch.cuda.synchronize()
    except Exception as exc:
        print("FAULT:", str(exc).splitlines()[0])
        if TRACE is not None:
            t = TRACE.view(-1, 512)
            for role in range(t.shape[0]):
                n = int(t[role, 0])
                if n:      # pairs of (code << 16 | pass - window start, clock)
                    ws = [int(t[role, 1 + 2 * i]) & 0xFFFFFFFF for i in range(min(n, 250))]
                    line = "role %2d: " % role + " ".join("%d@%d" % (w >> 16, w & 0xFFFF) for w in ws)
                    print(line)
                    with open(os.environ.get("TC3_TRACE_FILE", "/tmp/tc3_trace.txt"), "a") as f:
                        f.write(line + "\n")
        os._exit(1)
    pts = (ro[:, None] + rd[:, None] * z[..., None]).reshape(-1, 3).cpu()
    x = O.positional_encoding(pts, 10)
    if model.use_viewdirs:
        x = torch.cat([x, O.positional_encoding(vd.cpu()[:, None].expand(n, S, 3).reshape(-1, 3), 4)], -1)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    emu = O.flexible_forward(sd, x, skip_connect_every=model.skip_connect_every, use_viewdirs=model.use_viewdirs,
                             bf16=True).reshape(n, S, 4)
    err = (rf.cpu() - emu).abs().reshape(-1, 4)
    print(name, "max abs err vs bf16 emulation: %.3e  nan: %d" % (float(err.max()), int(torch.isnan(rf).sum())))
    if float(err.max()) > 1e-3:
        pad = (-err.shape[0]) % 128
        e = torch.cat([err, torch.zeros(pad, 4)]).reshape(-1, 128, 4)
        for ti in range(e.shape[0]):
            if float(e[ti].max()) > 1e-3:
                rows = (e[ti].max(dim=1).values > 1e-3).nonzero().flatten().tolist()
                print("  tile %3d (group %d, t %d): per-channel max %s  bad rows %d: %s" % (ti, ti // 3, ti % 3, ["%.1e" % float(v) for v in e[ti].max(dim=0).values], len(rows), rows[:12]))
