"""Bring-up aid for the fused backward launch: run it with soft flag waits (variant bit 6) and dump the flag table."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
from nerf import tensorcore, training
from nerf import _lib as L

n, S = int(sys.argv[1]) if len(sys.argv) > 1 else 1, int(sys.argv[2]) if len(sys.argv) > 2 else 100
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 64
torch.manual_seed(0)
model = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
prog = model.program(ex, ed); spec = tensorcore.spec_for(model, prog)
g = torch.Generator().manual_seed(1)
ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
vd = rd / rd.norm(dim=-1, keepdim=True)
z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
d_rf = (torch.randn(n, S, 4, generator=g) * 0.05).cuda()
rf, tape = training.query_train(model, prog, spec, ro, rd, vd, z)
lay = training.tape_layout(spec, n * S)
print("layout nl", lay["nl"], "tiles", lay["n_tiles"], "total", lay["total"], "tape bytes", tape.numel())
want = training.mlp_backward(model, prog, spec, tape.clone(), d_rf, n, S, what=3)
torch.cuda.synchronize()
t2 = tape.clone()
got = training.mlp_backward(model, prog, spec, t2, d_rf, n, S, what=4, variant=variant)
torch.cuda.synchronize()
nt = lay["n_tiles"]
flag_bytes = 17 * nt * 3 * 4
flags = t2[t2.numel() - ((flag_bytes + 127) // 128 * 128):][:flag_bytes].view(torch.int32)
print("ready rows (row x tile):")
print(flags[:17 * nt].view(17, nt)[:, :8].cpu())
print("consumed:", flags[17 * nt:].view(17, nt, 2)[:, :4].reshape(17, -1).cpu())
print("rel err vs split:", float((got - want).norm() / want.norm()))
