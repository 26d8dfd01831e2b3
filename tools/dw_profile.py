import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'dex-nerf_b200'))
import nerf
from nerf import training, tensorcore
torch.manual_seed(0)
model = nerf.FlexibleNeRFModel(8,256,4,10,4).cuda()
ex, ed = nerf.get_embedding_function(10,True,True), nerf.get_embedding_function(4,True,True)
prog = model.program(ex, ed); spec = tensorcore.spec_for(model, prog)
n, S = 4096, 192
lay = training.tape_layout(spec, n*S)
tape = torch.zeros(lay["total"], dtype=torch.uint8, device="cuda")
d_rf = torch.zeros(n, S, 4, device="cuda")
training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2, variant=0); torch.cuda.synchronize()
for variant in (128,):
    print("variant", variant, flush=True)
    training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2, variant=variant)
    torch.cuda.synchronize()
for variant in (0, 0, 6):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2, variant=variant); e1.record(); torch.cuda.synchronize()
    print("variant", variant, "whole call %.3f ms" % e0.elapsed_time(e1))
