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
for variant, name in ((0,"full"),(2,"no bias sums"),(4,"no MMA"),(6,"streaming only")):
    for _ in range(3): training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2, variant=variant)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); ts=[]
    for _ in range(5):
        flat = torch.zeros_like(model.packed_params())
        blob = tensorcore.packed_weights(model, prog, spec); blob_t = training.packed_weights_t(model, prog, spec)
        from nerf import _lib as L
        e0.record()
        L.check(L.lib().dexnerf_tc_backward(spec, prog, L.ptr(blob), L.ptr(blob_t), L.ptr(tape), L.ptr(d_rf), n, S, L.ptr(flat), 2, variant, L.stream_ptr()), "x")
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms=min(ts); print("%-16s %.3f ms  %.2f TB/s" % (name, ms, 6144*1424*1024/ms/1e9))
