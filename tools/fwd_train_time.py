"""The forward kernel with tape against the inference kernel on the same samples, each alone (CUDA events, median of 9):
what the tape costs the forward (tools/README.md).  Usage: python tools/fwd_train_time.py [rays] [samples]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf                                                     # noqa: E402
from nerf import tensorcore, training                           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
torch.manual_seed(0)
model = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
prog = model.program(ex, ed)
spec = tensorcore.spec_for(model, prog)
g = torch.Generator(device="cuda").manual_seed(0)
ro = torch.randn(n, 3, device="cuda", generator=g)
rd = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(2 + 4 * torch.rand(n, S, device="cuda", generator=g), dim=-1).values
rf = torch.empty(n, S, 4, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(9):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


t_inf = timed(lambda: tensorcore.query(model, prog, ro, rd, rd, z, rf))
t_tr = timed(lambda: training.query_train(model, prog, spec, ro, rd, rd, z))
print("%d rays x %d samples: inference %.3f ms, forward with tape %.3f ms (x %.2f)" % (n, S, t_inf, t_tr, t_tr / t_inf))
