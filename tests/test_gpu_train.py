"""Training path on the GPU (BASELINE config 4): compositing backward, the training tape of the
tensor-core forward, the activation-gradient chain, the weight-gradient GEMM and the whole
training iteration against the reference-generated gradients (tests/golden/train_grads.npz) and
the CPU oracle.  Tolerances are written at each comparison."""
import os

import numpy as np
import pytest
import torch

import nerf
from nerf import _lib as L
from nerf import tensorcore, training
from oracle import nerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu
t = torch.from_numpy


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def rel_err(a, b):
    return float((a - b).norm() / (b.norm() + 1e-20))


# ------------------------------------------------------------------ compositing backward
@pytest.mark.parametrize("n,S,white,with_noise", [(37, 48, False, False), (33, 64, True, True), (9, 192, False, True),
                                                  (5, 5, True, False), (300, 384, False, False)])
def test_volume_render_backward(n, S, white, with_noise):
    """dL/d(radiance_field) of volume_render_radiance_field against autograd through the oracle
    (fp32 graph, fp64 scans).  Bar: 2e-5 of the largest gradient entry per tensor."""
    g = torch.Generator().manual_seed(n * 1000 + S)
    rf = torch.randn(n, S, 4, generator=g)
    rf[..., 3] = 3.0 * torch.randn(n, S, generator=g) + 0.5
    rf[0, :, 3] = -1.0                      # nothing absorbs
    if n > 2:
        rf[2, :-1, 3] = -5.0
        rf[2, -1, 3] = 0.5                  # only the last (1e10) sample absorbs
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S, generator=g), dim=-1).values
    rd = torch.randn(n, 3, generator=g)
    noise = 0.2 * torch.randn(n, S, generator=g) if with_noise else None
    g_rgb, g_depth, g_acc = torch.randn(n, 3, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
    rfo = rf.clone().requires_grad_(True)
    res = O.volume_render_radiance_field(rfo, z, rd, 0.0, white, [], noise=noise)
    ((res[0] * g_rgb).sum() + (res[4] * g_depth).sum() + (res[2] * g_acc).sum()).backward()
    got = training.volume_render_backward(rf.cuda(), z.cuda(), rd.cuda(), None if noise is None else noise.cuda(),
                                          white, g_rgb.cuda(), g_depth.cuda(), g_acc.cuda()).cpu()
    ref = rfo.grad
    assert torch.isfinite(got).all()
    for ch in range(4):
        scale = float(ref[..., ch].abs().max()) + 1e-12
        assert float((got[..., ch] - ref[..., ch]).abs().max()) <= 2e-5 * scale, ch
    # rgb-only upstream gradient (what the training loss produces): NULL depth / acc pointers
    got2 = training.volume_render_backward(rf.cuda(), z.cuda(), rd.cuda(), None if noise is None else noise.cuda(),
                                           white, g_rgb.cuda(), None, None).cpu()
    rfo2 = rf.clone().requires_grad_(True)
    res = O.volume_render_radiance_field(rfo2, z, rd, 0.0, white, [], noise=noise)
    (res[0] * g_rgb).sum().backward()
    assert float((got2 - rfo2.grad).abs().max()) <= 2e-5 * float(rfo2.grad.abs().max())


# ------------------------------------------------------------------ helpers for the tensor-core path
def make_model(hidden, layers, skip, Lx, seed=3, boost=30.0):
    torch.manual_seed(seed)
    m = nerf.FlexibleNeRFModel(layers, hidden, skip, Lx, 4)
    with torch.no_grad():
        m.fc_alpha.weight.mul_(boost)
        m.fc_alpha.bias.fill_(1.0)
    return m.cuda()


def make_rays(n, S, seed=0):
    g = torch.Generator().manual_seed(seed)
    ro = (torch.randn(n, 3, generator=g) * 0.3).cuda()
    rd = torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    return ro, rd, vd, z


def emulate_forward(model, enc_xyz, enc_dir):
    """bf16-operand emulation in torch (on the GPU): list of the bf16 OUTPUT of every tensor-core
    layer, in the kernel's layer order, and the final (rgb, sigma)."""
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    skip, n_trunk = model.skip_connect_every, len(model.layers_xyz)
    xyz, dr = bf(enc_xyz), bf(enc_dir)
    acts = []
    h = xyz @ bf(sd["layer1.weight"]).t() + sd["layer1.bias"]
    acts.append(bf(h))
    for i in range(n_trunk):
        inp = bf(h)
        if i % skip == 0 and i > 0:
            inp = torch.cat((inp, xyz), -1)
        h = torch.relu(inp @ bf(sd[f"layers_xyz.{i}.weight"]).t() + sd[f"layers_xyz.{i}.bias"])
        acts.append(bf(h))
    sigma = h @ sd["fc_alpha.weight"].t() + sd["fc_alpha.bias"]
    feat = torch.relu(bf(h) @ bf(sd["fc_feat.weight"]).t() + sd["fc_feat.bias"])
    acts.append(bf(feat))
    y = torch.relu(torch.cat((bf(feat), dr), -1) @ bf(sd["layers_dir.0.weight"]).t() + sd["layers_dir.0.bias"])
    acts.append(bf(y))
    rgb = y @ sd["fc_rgb.weight"].t() + sd["fc_rgb.bias"]
    return acts, torch.cat((rgb, sigma), -1)


def run_forward_with_tape(model, Lx, n, S, seed=0):
    ex, ed = nerf.get_embedding_function(Lx, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    spec = tensorcore.spec_for(model, prog)
    ro, rd, vd, z = make_rays(n, S, seed)
    rf, tape = training.query_train(model, prog, spec, ro, rd, vd, z)
    torch.cuda.synchronize()
    lay = training.tape_layout(spec, n * S)
    pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3)
    enc_xyz = ex(pts)
    enc_dir = ed(vd[:, None, :].expand(n, S, 3).reshape(-1, 3).contiguous())
    return dict(prog=prog, spec=spec, rf=rf, tape=tape, lay=lay, enc_xyz=enc_xyz, enc_dir=enc_dir, n=n, S=S,
                ro=ro, rd=rd, vd=vd, z=z)


CONFIGS = [(256, 8, 4, 10), (128, 8, 3, 6)]


@pytest.mark.parametrize("hidden,layers,skip,Lx", CONFIGS)
def test_forward_tape(hidden, layers, skip, Lx):
    """The training variant of the forward returns the same field as the inference kernel and its
    tape holds the bf16 operand image of every layer plus the ReLU bits."""
    model = make_model(hidden, layers, skip, Lx)
    n, S = 7, 50                       # 350 samples: 3 tiles -> padded to 2 pairs
    r = run_forward_with_tape(model, Lx, n, S)
    M = n * S
    rf_inf = torch.empty_like(r["rf"])
    tensorcore.query(model, r["prog"], r["ro"], r["rd"], r["vd"], r["z"], rf_inf)
    if hidden == 256:
        assert torch.equal(rf_inf, r["rf"])
    else:
        # hidden-128 inference runs on the three-tile kernel (csrc/mlp_tc3.cu), which adds a layer's bias FIRST (as an
        # MMA), the training forward last: same products, another fp32 summation order, hence a few bf16 rounding flips
        assert float((rf_inf[..., :3] - r["rf"][..., :3]).abs().max()) < 1e-3
        assert float((rf_inf[..., 3] - r["rf"][..., 3]).abs().max()) < 4e-3 * max(1.0, float(r["rf"][..., 3].abs().max()))
    acts, out = emulate_forward(model, r["enc_xyz"], r["enc_dir"])
    lay, tape = r["lay"], r["tape"]
    nt = lay["n_tiles"]
    assert nt == 4 and lay["nl"] == layers + 2
    dx, dd = model.dim_xyz, model.dim_dir
    img = training.decode_image(tape, lay["xyz"], nt, 64)[:M]
    assert float((img[:, :dx] - bf(r["enc_xyz"])).abs().max()) <= 2 ** -7      # one bf16 ulp of |x| <= 1.x .. sin/cos
    assert float(img[:, dx:].abs().max()) == 0.0
    img = training.decode_image(tape, lay["dir"], nt, 32)[:M]
    assert float((img[:, :dd] - bf(r["enc_dir"])).abs().max()) <= 2 ** -7
    for l, a in enumerate(acts):
        width = a.shape[1]
        img = training.decode_image(tape, lay["act"][l], nt, width)[:M]
        scale = float(a.abs().max())
        # bf16 rounding of slightly different fp32 sums: a few ulps on a few entries, 2e-2 relative worst case
        assert float((img - a).abs().max()) <= 2e-2 * scale, l
        assert rel_err(img, a) < 3e-3, l
        if l > 0:                           # ReLU layers: mask bit == (stored activation > 0)
            slots = 2 * ((width + 127) // 128)      # two 64-column (or narrower) blocks per 128-wide pass
            words = tape[lay["mask"][l]: lay["mask"][l] + nt * slots * 1024].view(torch.int64).view(nt, slots, 128)
            cols = width // slots
            # column f of a 64-column slot: 32-bit word f >> 5, 16-bit slice (f >> 4) & 1, and inside the slice the
            # even columns at bits 7..0, the odd ones at bits 15..8 (relu_bit_pos in csrc/mlp_tc_shared.cuh)
            f = torch.arange(cols, device="cuda")
            pos = (f >> 4) * 16 + (f & 1) * 8 + 7 - ((f & 15) >> 1)
            bits = ((words[..., None] >> pos) & 1).bool()    # tile, slot, row, col
            bits = bits.permute(0, 2, 1, 3).reshape(nt * 128, width)[:M]
            assert torch.equal(bits, img > 0), l


def emulate_backward(model, r, d_rf):
    """bf16-operand emulation of the activation-gradient chain using the TAPE's own activations
    (so that ReLU masks agree bit for bit): G per tensor-core layer, in layer order."""
    sd = {k: v.detach().float() for k, v in model.state_dict().items()}
    lay, tape, M = r["lay"], r["tape"], r["n"] * r["S"]
    nt, nl, H = lay["n_tiles"], lay["nl"], model.hidden_size
    n_trunk = len(model.layers_xyz)
    act = [training.decode_image(tape, lay["act"][l], nt, H if l < nl - 1 else H // 2)[:M] for l in range(nl)]
    G = [None] * nl
    drgb, dsig = d_rf[:, :3], bf(d_rf[:, 3:4])
    G[nl - 1] = bf((drgb @ sd["fc_rgb.weight"]) * (act[nl - 1] > 0))
    G[nl - 2] = bf((G[nl - 1] @ bf(sd["layers_dir.0.weight"])[:, :H]) * (act[nl - 2] > 0))
    dh = G[nl - 2] @ bf(sd["fc_feat.weight"]) + dsig * sd["fc_alpha.weight"]
    G[nl - 3] = bf(dh * (act[nl - 3] > 0))
    for i in range(n_trunk - 1, -1, -1):          # trunk layer i is tensor-core layer i + 1
        dh = G[i + 1] @ bf(sd[f"layers_xyz.{i}.weight"])[:, :H]
        G[i] = bf(dh * (act[i] > 0)) if i > 0 else bf(dh)
    return act, G


@pytest.mark.parametrize("hidden,layers,skip,Lx", CONFIGS)
def test_activation_gradient_chain(hidden, layers, skip, Lx):
    model = make_model(hidden, layers, skip, Lx)
    n, S = 9, 64
    r = run_forward_with_tape(model, Lx, n, S, seed=1)
    M = n * S
    d_rf = torch.randn(M, 4, generator=torch.Generator().manual_seed(5)).cuda() * 0.1
    training.mlp_backward(model, r["prog"], r["spec"], r["tape"], d_rf.view(n, S, 4).contiguous(), n, S, what=1)
    torch.cuda.synchronize()
    act, G = emulate_backward(model, r, d_rf)
    lay, nt = r["lay"], r["lay"]["n_tiles"]
    for l in range(lay["nl"] - 1, -1, -1):
        got = training.decode_image(r["tape"], lay["grad"][l], nt, G[l].shape[1])
        assert float(got[M:].abs().max()) == 0.0 if got.shape[0] > M else True      # padded rows carry no gradient
        e = rel_err(got[:M], G[l])
        print("G layer", l, "rel_err %.5f" % e)
        assert e < 1e-2, (l, e)                 # bf16 re-rounding of fp32 sums accumulated in another order
    got = training.decode_image(r["tape"], lay["ghead"], nt, 16)[:M]
    assert torch.equal(got[:, :4], bf(d_rf)) and float(got[:, 4:].abs().max()) == 0.0


def reference_weight_grads(model, prog, act, G, enc_xyz_img, enc_dir_img, ghead):
    """dWt = X^T G in fp32 from the bf16 images, assembled in the program layout."""
    flat = torch.zeros_like(model.packed_params())
    layers = model._layers()
    H, nl = model.hidden_size, len(act)
    n_trunk = len(model.layers_xyz)
    # tensor-core layer l -> program op: layer1 0, trunk 1..n_trunk, fc_alpha n_trunk+1, fc_feat +2, dir +3, rgb +4
    tc_ops = list(range(0, n_trunk + 1)) + [n_trunk + 2, n_trunk + 3]
    for l, opi in enumerate(tc_ops):
        lin = layers[opi][0]
        op = prog.ops[opi]
        if l == 0:
            X = enc_xyz_img[:, :lin.in_features]
        else:
            X = act[l - 1]
            if lin.in_features > X.shape[1]:
                extra = enc_dir_img if l == nl - 1 else enc_xyz_img
                X = torch.cat((X, extra[:, :lin.in_features - X.shape[1]]), -1)
        flat[op.w_off:op.w_off + lin.in_features * lin.out_features] = (X.t() @ G[l]).reshape(-1)
        flat[op.b_off:op.b_off + lin.out_features] = G[l].sum(0)
    oa, orgb = prog.ops[n_trunk + 1], prog.ops[n_trunk + 4]
    flat[oa.w_off:oa.w_off + H] = (act[nl - 3].t() @ ghead[:, 3:4]).reshape(-1)
    flat[oa.b_off] = ghead[:, 3].sum()
    flat[orgb.w_off:orgb.w_off + 3 * (H // 2)] = (act[nl - 1].t() @ ghead[:, :3]).reshape(-1)
    flat[orgb.b_off:orgb.b_off + 3] = ghead[:, :3].sum(0)
    return flat


@pytest.mark.parametrize("hidden,layers,skip,Lx", CONFIGS)
def test_weight_gradient_gemm(hidden, layers, skip, Lx):
    """The split-K weight-gradient GEMM in isolation: random bf16 images written into a tape,
    dWt = X^T G compared with torch matmul on the same values (fp32 accumulate both sides)."""
    model = make_model(hidden, layers, skip, Lx)
    ex, ed = nerf.get_embedding_function(Lx, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    spec = tensorcore.spec_for(model, prog)
    n, S = 20, 64                           # 1280 samples = 10 tiles
    lay = training.tape_layout(spec, n * S)
    nt, nl, H = lay["n_tiles"], lay["nl"], hidden
    tape = torch.zeros(lay["total"], dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(11)
    rnd = lambda w: bf(torch.randn(nt * 128, w, generator=g, device="cuda"))
    act = [rnd(H if l < nl - 1 else H // 2) for l in range(nl)]
    G = [bf(rnd(H if l < nl - 1 else H // 2) * 0.1) for l in range(nl)]
    xyz, dr, ghead = rnd(64), rnd(32), rnd(16)
    ghead[:, 4:] = 0
    for l in range(nl):
        training.encode_image(tape, lay["act"][l], nt, act[l])
        training.encode_image(tape, lay["grad"][l], nt, G[l])
    training.encode_image(tape, lay["xyz"], nt, xyz)
    training.encode_image(tape, lay["dir"], nt, dr)
    training.encode_image(tape, lay["ghead"], nt, ghead)
    assert torch.equal(training.decode_image(tape, lay["act"][1], nt, H), act[1])
    want = reference_weight_grads(model, prog, act, G, xyz, dr, ghead)
    d_rf = torch.zeros(n, S, 4, device="cuda")
    errs = {}
    for variant in (0, 1):
        got = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2, variant=variant)
        torch.cuda.synchronize()
        errs[variant] = rel_err(got, want)
    print("dW rel_err by descriptor variant", errs)
    got = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=2)
    bad = []
    for i, (lin, *_r) in enumerate(model._layers()):
        op = prog.ops[i]
        for kind, sl in (("weight", slice(op.w_off, op.w_off + lin.in_features * lin.out_features)),
                         ("bias", slice(op.b_off, op.b_off + lin.out_features))):
            e = rel_err(got[sl], want[sl])
            print("op", i, kind, "rel_err %.2e" % e, "max abs diff %.3e" % float((got[sl] - want[sl]).abs().max()))
            if e >= 1e-4:
                bad.append((i, kind, e))
    assert not bad, bad                  # same bf16 inputs, fp32 accumulation: only summation order differs
    assert errs[0] < 1e-4, errs


@pytest.mark.parametrize("what", [4, 8])
@pytest.mark.parametrize("hidden,layers,skip,n,S", [(256, 8, 4, 1, 100), (256, 8, 4, 37, 64), (128, 8, 3, 300, 192),
                                                    (128, 4, 4, 640, 128), (256, 8, 4, 1500, 192)])
def test_single_launch_backwards_equal_the_two_kernels(hidden, layers, skip, n, S, what):
    """dexnerf_tc_backward(what = 4 / 8) - the activation-gradient chain and the weight-gradient GEMM in ONE launch
    (4: on disjoint SMs; 8: as two warp groups of every CTA, with back-pressure), gradient images handed over through
    L2 with release / acquire flags and discarded after use - against what = 3 (the stand-alone chain kernel, then the
    stand-alone GEMM, images through HBM; the training path) on the same tape: the same bf16 images enter the same
    MMAs, only the split-K partition (hence the order of the fp32 red.adds) differs.  Sizes: a single padded tile,
    an odd tile count, more tiles than SMs, and 2 250 tiles."""
    torch.manual_seed(hidden + n)
    model = nerf.FlexibleNeRFModel(layers, hidden, skip, 10, 4).cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    spec = tensorcore.spec_for(model, prog)
    g = torch.Generator().manual_seed(S)
    ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    d_rf = (torch.randn(n, S, 4, generator=g) * 0.05).cuda()
    rf, tape = training.query_train(model, prog, spec, ro, rd, vd, z)
    tape2 = tape.clone()
    want = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=3)
    got = training.mlp_backward(model, prog, spec, tape2, d_rf, n, S, what=what)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all() and float(want.abs().max()) > 0
    # measured: identical at one tile, 4e-5 at 2 250 tiles (fp32 summation of ~3e5 products in two different orders)
    assert rel_err(got, want) < 1e-4, rel_err(got, want)
    for i, (lin, *_r) in enumerate(model._layers()):
        op = prog.ops[i]
        for sl in (slice(op.w_off, op.w_off + lin.in_features * lin.out_features), slice(op.b_off, op.b_off + lin.out_features)):
            if float(want[sl].norm()) > 1e-12:
                assert rel_err(got[sl], want[sl]) < 5e-4, (i, rel_err(got[sl], want[sl]))
    # the gradient images the chain left in the tape are the stand-alone chain's, bit for bit, wherever a reader did
    # not discard them... which it did: the fused launch drops them from L2, so only the split path's tape holds them.
    # A second fused launch on the same tape must give the same result again (flags are re-armed per launch).
    again = training.mlp_backward(model, prog, spec, tape2, d_rf, n, S, what=what)
    assert rel_err(again, want) < 1e-4


def test_backward_kernels_on_a_subset_of_the_sms():
    """`variant` bits 8-15 / 16-23 of dexnerf_tc_backward: the weight-gradient GEMM on 100 SMs and the chain on 40 give
    the gradients of the full-width launches (another split-K partition: fp32 summation order only)."""
    torch.manual_seed(5)
    model = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    spec = tensorcore.spec_for(model, prog)
    n, S = 300, 192
    g = torch.Generator().manual_seed(2)
    ro, rd = (torch.randn(n, 3, generator=g) * 0.3).cuda(), torch.randn(n, 3, generator=g).cuda()
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values.cuda()
    d_rf = (torch.randn(n, S, 4, generator=g) * 0.05).cuda()
    rf, tape = training.query_train(model, prog, spec, ro, rd, vd, z)
    tape2 = tape.clone()
    want = training.mlp_backward(model, prog, spec, tape, d_rf, n, S, what=3)
    training.mlp_backward(model, prog, spec, tape2, d_rf, n, S, what=1, variant=40 << 16)
    got = training.mlp_backward(model, prog, spec, tape2, d_rf, n, S, what=2, variant=100 << 8)
    assert torch.equal(tape2, tape)              # the chain's gradient images do not depend on the grid
    assert rel_err(got, want) < 1e-4


SPLIT_WORKER = r'''
import os, sys, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import nerf
torch.manual_seed(3)
mc, mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda(), nerf.FlexibleNeRFModel(8, 256, 4, 10, 4).cuda()
mode = dict(chunksize=1 << 20, perturb=True, num_coarse=64, num_fine=128, white_background=False,
            radiance_field_noise_std=0.2, lindisp=False)
cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=2.0, far=6.0), nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
g = torch.Generator().manual_seed(5)
n = 512
ro = (torch.randn(n, 3, generator=g) * 0.2).cuda(); rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
tgt = torch.rand(n, 3, generator=g).cuda()
rng = dict(t_rand=torch.rand(n, 64, generator=g).cuda(), u=torch.rand(n, 128, generator=g).cuda(),
           noise_coarse=(0.2 * torch.randn(n, 64, generator=g)).cuda(), noise_fine=(0.2 * torch.randn(n, 192, generator=g)).cuda())
tr = nerf.Trainer(mc, mf, cfg, ex, ed)
losses = [float(tr.step(ro, rd, tgt, rng=rng)[0]) for _ in range(3)]
torch.cuda.synchronize()
torch.save(dict(losses=losses, params=tr.params.cpu()), sys.argv[1])
'''


def test_trainer_with_the_two_backwards_side_by_side(tmp_path):
    """DEXNERF_BWD_SPLIT=40 (render.cu: the fine network's weight-gradient GEMM on 108 SMs next to the coarse network's
    compositing backward + chain on 40, second stream, fork / join events) against the default order: the same losses
    and parameters after three iterations, up to the summation order of the weight gradients."""
    import subprocess
    import sys
    script = tmp_path / "split_worker.py"
    script.write_text("ROOT = %r\n" % ROOT + SPLIT_WORKER)
    outs = []
    for k in ("0", "40"):
        path = str(tmp_path / ("out%s.pt" % k))
        r = subprocess.run([sys.executable, str(script), path], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, DEXNERF_BWD_SPLIT=k))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(torch.load(path))
    a, b = outs
    assert all(abs(x - y) < 1e-5 for x, y in zip(a["losses"], b["losses"])), (a["losses"], b["losses"])
    diff = (a["params"] - b["params"]).abs()
    assert float((diff > 2e-4).float().mean()) < 0.02        # Adam steps of entries with |grad| ~ eps aside


# ------------------------------------------------------------------ whole training iteration
def make_cfg(nc, nf, near, far, white, noise_std=0.2, perturb=True):
    mode = dict(chunksize=1 << 20, perturb=perturb, num_coarse=nc, num_fine=nf, white_background=white,
                radiance_field_noise_std=noise_std, lindisp=False)
    return nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=near, far=far),
                             nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))


def test_training_iteration_matches_reference(golden):
    """The reference's own training iteration (fixture made by tests/golden/make_golden.py
    gen_train_grads from the unmodified reference): 8-layer x 128 FlexibleNeRFModel pair, RNG
    replayed.  Bars (bf16 tensor-core operands vs the reference's fp32): loss within 2e-3;
    whole-network gradient within 4 % in norm (measured 1.3 % / 2.9 %), cosine > 0.998, every tensor within 15 %."""
    g = golden("train_grads")
    tag = "h128"
    hidden, layers, skip, white = map(int, g[f"{tag}.cfg"])
    nets = []
    for net in ("coarse", "fine"):
        m = nerf.FlexibleNeRFModel(layers, hidden, skip, 6, 4)
        m.load_state_dict({k[len(tag) + len(net) + 2:]: t(v) for k, v in g.items() if k.startswith(f"{tag}.{net}.")})
        nets.append(m.cuda())
    mc, mf = nets
    ex, ed = nerf.get_embedding_function(6, True, True), nerf.get_embedding_function(4, True, True)
    rng = dict(t_rand=t(g[f"{tag}.t_rand"]).cuda(), u=t(g[f"{tag}.u"]).cuda(),
               noise_coarse=t(g[f"{tag}.noise_c"]).cuda(), noise_fine=t(g[f"{tag}.noise_f"]).cuda())
    H, W = map(int, g["HW"])
    out = nerf.run_one_iter_of_nerf(H, W, 9.0, mc, mf, t(g["ro"]).cuda(), t(g["rd"]).cuda(),
                                    make_cfg(16, 24, 2.0, 6.0, bool(white)), mode="train", encode_position_fn=ex,
                                    encode_direction_fn=ed, m_thres_cand=g["thr"].tolist(), rng=rng)
    assert out[0].requires_grad and out[3].requires_grad and len(out) == 6 + len(g["thr"])
    target = t(g["target"]).cuda()
    loss = torch.nn.functional.mse_loss(out[0][..., :3], target) + torch.nn.functional.mse_loss(out[3][..., :3], target)
    loss.backward()
    assert abs(float(loss) - float(g[f"{tag}.loss"])) < 2e-3
    worst = []
    for net, m in (("coarse", mc), ("fine", mf)):
        num = den = dot = 0.0
        for k, p in m.named_parameters():
            ref = t(g[f"{tag}.grad.{net}.{k}"]).cuda()
            assert p.grad is not None and p.grad.shape == ref.shape, k
            assert torch.isfinite(p.grad).all(), k
            print("grad", net, k, "rel_err %.4f" % rel_err(p.grad, ref), "ref norm %.3e" % float(ref.norm()))
            if float(ref.norm()) > 1e-7:
                worst.append((rel_err(p.grad, ref), net, k))
            num += float(((p.grad - ref) ** 2).sum()); den += float((ref ** 2).sum()); dot += float((p.grad * ref).sum())
        got_norm = sum(float((p.grad ** 2).sum()) for p in m.parameters()) ** 0.5
        print("net", net, "total rel_err %.4f" % ((num / den) ** 0.5), "cosine %.5f" % (dot / (den ** 0.5 * got_norm)))
        assert (num / den) ** 0.5 < 4e-2, net            # measured: coarse 0.0125, fine 0.0294 (= the bf16 oracle's)
        assert dot / (den ** 0.5 * got_norm) > 0.998, net
    assert max(worst)[0] < 0.15, max(worst)


def test_training_iteration_8x256_vs_oracle():
    """BASELINE config 4's networks (8x256 skip 4, L=10/4, 64 + 128 samples) on a small ray batch
    against the CPU oracle's autograd with the same bf16 operand contract."""
    torch.manual_seed(42)
    mc, mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4), nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    with torch.no_grad():
        for m in (mc, mf):
            m.fc_alpha.weight.mul_(40.0)
            m.fc_alpha.bias.fill_(0.5)
    sdc = {k: v.detach().clone() for k, v in mc.state_dict().items()}
    sdf = {k: v.detach().clone() for k, v in mf.state_dict().items()}
    mc, mf = mc.cuda(), mf.cuda()
    n, nc, nf = 48, 64, 128
    g = torch.Generator().manual_seed(9)
    T = O.pose_spherical_world2cam(30.0, -30.0, 4.0)
    K = torch.tensor([[60.0, 0, 4.0], [0, 60.0, 3.0], [0, 0, 1]])
    ro, rd = O.get_ray_bundle(6, 8, None, T, K)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    target = torch.rand(n, 3, generator=g)
    rng = dict(t_rand=torch.rand(n, nc, generator=g), u=torch.rand(n, nf, generator=g),
               noise_coarse=0.2 * torch.randn(n, nc, generator=g), noise_fine=0.2 * torch.randn(n, nc + nf, generator=g))
    opts = O.RenderOptions(near=2.0, far=6.0, num_coarse=nc, num_fine=nf, Lx=10, Ld=4, perturb=True, noise_std=0.2)
    loss_o, _, _, gc, gf = O.train_loss_and_grads(sdc, sdf, ro, rd, target, opts, [], 4, 4, t_rand=rng["t_rand"],
                                                  u=rng["u"], noise_coarse=rng["noise_coarse"],
                                                  noise_fine=rng["noise_fine"], bf16=True)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=5e-3)
    before = {k: v.detach().clone() for k, v in mc.state_dict().items()}
    loss, _, _ = nerf.train_step(mc, mf, opt, ro.cuda(), rd.cuda(), target.cuda(), make_cfg(nc, nf, 2.0, 6.0, False),
                                 ex, ed, m_thres_cand=[], rng={k: v.cuda() for k, v in rng.items()}, height=6, width=8,
                                 focal=60.0)
    assert abs(float(loss) - float(loss_o)) < 2e-3
    # the optimizer stepped: parameters moved by about lr in the direction of -sign(grad)
    moved = sum(float((mc.state_dict()[k] - before[k]).abs().sum()) for k in before)
    assert moved > 0
    # gradients (recomputed, optimizer.zero_grad() cleared them): compare with the oracle
    for m, sd in ((mc, sdc), (mf, sdf)):
        m.load_state_dict(sd)
    out = nerf.run_one_iter_of_nerf(6, 8, 60.0, mc, mf, ro.cuda(), rd.cuda(), make_cfg(nc, nf, 2.0, 6.0, False),
                                    mode="train", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[],
                                    rng={k: v.cuda() for k, v in rng.items()})
    tg = target.cuda()
    (torch.nn.functional.mse_loss(out[0], tg) + torch.nn.functional.mse_loss(out[3], tg)).backward()
    for m, grads in ((mc, gc), (mf, gf)):
        num = den = 0.0
        for k, p in m.named_parameters():
            ref = grads[k].cuda()
            num += float(((p.grad - ref) ** 2).sum()); den += float((ref ** 2).sum())
        assert (num / den) ** 0.5 < 3e-2


# ------------------------------------------------------------------ flat-buffer Trainer
def test_adam_and_mse_kernels_vs_torch():
    g = torch.Generator(device="cuda").manual_seed(2)
    n = 100_003
    p = torch.randn(n, device="cuda", generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=5e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(n, device="cuda", generator=g) * (10.0 ** -step)
        ref.grad = grad.clone()
        opt.step()
        L.check(L.lib().dexnerf_adam_step(L.ptr(p), L.ptr(grad), L.ptr(m), L.ptr(v), n, 5e-3, 0.9, 0.999, 1e-8, step, 1.0,
                                          L.stream_ptr()), "adam")
        assert float((p - ref.detach()).abs().max()) < 2e-6, step      # fp32 rounding of one update
    pred, tgt = torch.rand(4096, 3, device="cuda", generator=g), torch.rand(4096, 3, device="cuda", generator=g)
    gout, loss = torch.empty_like(pred), torch.zeros(1, device="cuda")
    L.check(L.lib().dexnerf_mse_loss_grad(L.ptr(pred), L.ptr(tgt), pred.numel(), pred.numel(), L.ptr(gout), L.ptr(loss),
                                          L.stream_ptr()),
            "mse")
    pr = pred.clone().requires_grad_(True)
    lr_ = torch.nn.functional.mse_loss(pr, tgt)
    lr_.backward()
    assert abs(float(loss) - float(lr_)) < 1e-6 and float((gout - pr.grad).abs().max()) < 1e-9


def test_trainer_matches_autograd_path():
    """nerf.Trainer (flat buffers, fused Adam) against run_one_iter_of_nerf(mode='train') +
    torch.optim.Adam on the same networks, rays and random draws: same loss, same parameter update."""
    import copy
    torch.manual_seed(11)
    mc = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    mf = nerf.FlexibleNeRFModel(8, 256, 4, 10, 4)
    with torch.no_grad():
        for m in (mc, mf):
            m.fc_alpha.weight.mul_(40.0)
            m.fc_alpha.bias.fill_(0.5)
    mc, mf = mc.cuda(), mf.cuda()
    mc2, mf2 = copy.deepcopy(mc), copy.deepcopy(mf)
    n, nc, nf = 96, 64, 128
    g = torch.Generator().manual_seed(4)
    T = O.pose_spherical_world2cam(30.0, -30.0, 4.0)
    K = torch.tensor([[60.0, 0, 6.0], [0, 60.0, 4.0], [0, 0, 1]])
    ro, rd = O.get_ray_bundle(8, 12, None, T, K)
    ro, rd = ro.reshape(-1, 3).cuda(), rd.reshape(-1, 3).cuda()
    target = torch.rand(n, 3, generator=g).cuda()
    cfg = make_cfg(nc, nf, 2.0, 6.0, False)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    trainer = nerf.Trainer(mc2, mf2, cfg, ex, ed, lr=5e-3)
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=5e-3)
    before = [p.detach().clone() for m in (mc, mf) for p in m.parameters()]
    for it in range(2):
        rng = dict(t_rand=torch.rand(n, nc, generator=g).cuda(), u=torch.rand(n, nf, generator=g).cuda(),
                   noise_coarse=(0.2 * torch.randn(n, nc, generator=g)).cuda(),
                   noise_fine=(0.2 * torch.randn(n, nc + nf, generator=g)).cuda())
        for pg in opt.param_groups:          # the script's schedule: the rate set after the previous iteration
            pg["lr"] = nerf.learning_rate(5e-3, it - 1, 250, 0.1) if it else 5e-3
        loss_a, _, _ = nerf.train_step(mc, mf, opt, ro, rd, target, cfg, ex, ed, m_thres_cand=[], rng=rng,
                                       height=8, width=12, focal=60.0)
        loss_t = trainer.step(ro, rd, target, rng=rng)
        assert abs(float(loss_t[0]) - float(loss_a)) < 1e-5, it
    trainer.sync_to_modules()
    after_a = [p.detach() for m in (mc, mf) for p in m.parameters()]
    after_t = [p.detach() for m in (mc2, mf2) for p in m.parameters()]
    num = sum(float(((a - t_) ** 2).sum()) for a, t_ in zip(after_a, after_t))
    den = sum(float(((a - b) ** 2).sum()) for a, b in zip(after_a, before))
    assert den > 0 and (num / den) ** 0.5 < 2e-2         # entries with |grad| ~ eps move differently; the rest agree
    # the synced modules render like the trainer's own weights
    with torch.no_grad():
        a = nerf.run_one_iter_of_nerf(8, 12, 60.0, mc2, mf2, ro, rd, make_cfg(nc, nf, 2.0, 6.0, False, 0.0, False),
                                      mode="validation", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[])
    assert torch.isfinite(a[3]).all()


def test_trainer_reduces_the_loss():
    torch.manual_seed(5)
    mc, mf = nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda(), nerf.FlexibleNeRFModel(8, 128, 3, 10, 4).cuda()
    cfg = make_cfg(32, 32, 2.0, 6.0, False)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    trainer = nerf.Trainer(mc, mf, cfg, ex, ed, lr=5e-3)
    g = torch.Generator().manual_seed(1)
    T = O.pose_spherical_world2cam(10.0, -20.0, 4.0)
    K = torch.tensor([[40.0, 0, 16.0], [0, 40.0, 16.0], [0, 0, 1]])
    ro, rd = O.get_ray_bundle(32, 32, None, T, K)
    ro, rd = ro.reshape(-1, 3).cuda(), rd.reshape(-1, 3).cuda()
    target = (0.5 + 0.4 * torch.sin(torch.arange(1024)[:, None] * torch.tensor([0.01, 0.02, 0.03]))).cuda()
    losses = [float(trainer.step(ro, rd, target)[0]) for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
    assert trainer.iteration == 60 and trainer.learning_rate() < 5e-3


def test_trainer_chunked_batch_equals_single_pass():
    """chunksize < batch: the chunks' gradients accumulate in the flat buffer before the one Adam step."""
    import copy
    torch.manual_seed(3)
    mc, mf = nerf.FlexibleNeRFModel(8, 128, 3, 6, 4).cuda(), nerf.FlexibleNeRFModel(8, 128, 3, 6, 4).cuda()
    mc2, mf2 = copy.deepcopy(mc), copy.deepcopy(mf)
    ex, ed = nerf.get_embedding_function(6, True, True), nerf.get_embedding_function(4, True, True)
    n, nc, nf = 96, 16, 24
    g = torch.Generator().manual_seed(8)
    ro = (torch.randn(n, 3, generator=g) * 0.2).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    target = torch.rand(n, 3, generator=g).cuda()
    rng = dict(t_rand=torch.rand(n, nc, generator=g).cuda(), u=torch.rand(n, nf, generator=g).cuda(),
               noise_coarse=(0.2 * torch.randn(n, nc, generator=g)).cuda(),
               noise_fine=(0.2 * torch.randn(n, nc + nf, generator=g)).cuda())
    cfg_a, cfg_b = make_cfg(nc, nf, 2.0, 6.0, True), make_cfg(nc, nf, 2.0, 6.0, True)
    cfg_b.nerf.train.chunksize = 40                     # 96 rays -> chunks of 40, 40, 16
    ta, tb = nerf.Trainer(mc, mf, cfg_a, ex, ed), nerf.Trainer(mc2, mf2, cfg_b, ex, ed)
    ta.keep_grads = tb.keep_grads = True                # leave the gradients in place for the comparison below
    la, lb = ta.step(ro, rd, target, rng=rng).clone(), tb.step(ro, rd, target, rng=rng).clone()
    assert float((la - lb).abs().max()) < 1e-6
    assert rel_err(tb.grads, ta.grads) < 1e-5
    upd_a, upd_b = ta.params.clone(), tb.params.clone()
    assert float((upd_a - upd_b).abs().max()) < 5e-3 + 1e-9      # at most one Adam step (lr) where |grad| ~ eps
    assert float(((upd_a - upd_b).abs() > 1e-5).float().mean()) < 0.02


def test_trainer_checkpoint_round_trip_with_torch_adam(tmp_path):
    """A checkpoint written by nerf.Trainer has the reference's layout (train_dexnerf_rgb.py:442-457):
    fresh modules + a fresh torch.optim.Adam resume from it (as the reference script does at :167-174)
    and their next iteration equals the trainer's next iteration; a new Trainer resumes from it too."""
    import copy
    torch.manual_seed(21)
    mk = lambda: nerf.FlexibleNeRFModel(8, 128, 3, 6, 4).cuda()
    mc, mf = mk(), mk()
    ex, ed = nerf.get_embedding_function(6, True, True), nerf.get_embedding_function(4, True, True)
    n, nc, nf = 64, 16, 24
    cfg = make_cfg(nc, nf, 2.0, 6.0, False)
    g = torch.Generator().manual_seed(3)
    ro = (torch.randn(n, 3, generator=g) * 0.2).cuda()
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).cuda()
    target = torch.rand(n, 3, generator=g).cuda()

    def draws():
        return dict(t_rand=torch.rand(n, nc, generator=g).cuda(), u=torch.rand(n, nf, generator=g).cuda(),
                    noise_coarse=(0.2 * torch.randn(n, nc, generator=g)).cuda(),
                    noise_fine=(0.2 * torch.randn(n, nc + nf, generator=g)).cuda())
    trainer = nerf.Trainer(mc, mf, cfg, ex, ed, lr=5e-3)
    for _ in range(3):
        loss = trainer.step(ro, rd, target, rng=draws())
    path = tmp_path / "checkpoint00003.ckpt"
    torch.save(trainer.checkpoint_dict(loss=float(loss[0]), psnr=float(nerf.mse2psnr(float(loss[0])))), path)
    ckpt = torch.load(path, weights_only=False)
    assert set(ckpt) == {"iter", "model_coarse_state_dict", "model_fine_state_dict", "optimizer_state_dict", "loss", "psnr"}
    # the reference's conventions: "iter" = 0-based index of the iteration just finished, Adam's "step" = updates made,
    # param_group lr = the rate set after that iteration (train_dexnerf_rgb.py:283-289, 443)
    assert ckpt["iter"] == 2 and set(ckpt["model_coarse_state_dict"]) == set(mk().state_dict())
    assert float(ckpt["optimizer_state_dict"]["state"][0]["step"]) == 3.0
    assert abs(ckpt["optimizer_state_dict"]["param_groups"][0]["lr"] - nerf.learning_rate(5e-3, 2, 250, 0.1)) < 1e-12
    # (a) the reference's resume path: modules + torch.optim.Adam
    mc2, mf2 = mk(), mk()
    mc2.load_state_dict(ckpt["model_coarse_state_dict"])
    mf2.load_state_dict(ckpt["model_fine_state_dict"])
    opt = torch.optim.Adam(list(mc2.parameters()) + list(mf2.parameters()), lr=5e-3)
    opt.load_state_dict(ckpt["optimizer_state_dict"])
    # (b) a new trainer
    mc3, mf3 = mk(), mk()
    t3 = nerf.Trainer(mc3, mf3, cfg, ex, ed, lr=5e-3)
    t3.load_checkpoint_dict(ckpt)
    # resumed as the reference script resumes: loop index = ckpt["iter"], Adam step and rate from the optimizer state
    assert t3.iteration == 2 and t3.adam_steps == 3 and t3.learning_rate() == trainer.learning_rate()
    rng = draws()
    la = nerf.train_step(mc2, mf2, opt, ro, rd, target, cfg, ex, ed, m_thres_cand=[], rng=rng, height=8, width=8, focal=1.0)[0]
    lb = trainer.step(ro, rd, target, rng=rng)[0].clone()
    lc = t3.step(ro, rd, target, rng=rng)[0].clone()
    assert abs(float(la) - float(lb)) < 1e-5 and abs(float(lb) - float(lc)) < 1e-6
    trainer.sync_to_modules()
    t3.sync_to_modules()
    for a, b, c in zip(mc2.parameters(), mc.parameters(), mc3.parameters()):
        assert float((b - c).abs().max()) < 1e-6                     # trainer vs resumed trainer: same kernels
        assert float(((a - b).abs() > 2e-4).float().mean()) < 0.02   # vs torch Adam: entries with |grad| ~ eps aside


def test_training_iteration_as_run_4x128_lego_checkpoint(golden):
    """The network every reference script actually instantiates (4 x 128, the constructor defaults,
    SURVEY.md section 8a-3) with the TRAINED pretrained/lego-lowres weights (a realistic, absorbing
    sigma field; fixture lego_lowres.npz): one training iteration, 64 + 64 samples, white background as
    in the blender configs, against the CPU oracle's autograd under the same bf16 operand contract."""
    g = golden("lego_lowres")
    mk = lambda: nerf.FlexibleNeRFModel(num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    sds = [{k[len(p):]: t(g[k]) for k in g.files if k.startswith(p)} for p in ("coarse.", "fine.")]
    mc, mf = mk(), mk()
    mc.load_state_dict(sds[0])
    mf.load_state_dict(sds[1])
    mc, mf = mc.cuda(), mf.cuda()
    H, W = map(int, g["HW"])
    ro, rd = O.get_ray_bundle(H, W, None, t(g["T"]), t(g["K"]))
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    n, nc, nf = ro.shape[0], 64, 64
    gen = torch.Generator().manual_seed(12)
    target = torch.rand(n, 3, generator=gen)
    rng = dict(t_rand=torch.rand(n, nc, generator=gen), u=torch.rand(n, nf, generator=gen),
               noise_coarse=0.2 * torch.randn(n, nc, generator=gen), noise_fine=0.2 * torch.randn(n, nc + nf, generator=gen))
    opts = O.RenderOptions(near=2.0, far=6.0, num_coarse=nc, num_fine=nf, Lx=10, Ld=4, perturb=True, noise_std=0.2,
                           white_background=True)
    loss_o, _, _, gc, gf = O.train_loss_and_grads(sds[0], sds[1], ro, rd, target, opts, [], 4, 4, t_rand=rng["t_rand"],
                                                  u=rng["u"], noise_coarse=rng["noise_coarse"], noise_fine=rng["noise_fine"],
                                                  bf16=True)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    out = nerf.run_one_iter_of_nerf(H, W, 14.0, mc, mf, ro.cuda(), rd.cuda(), make_cfg(nc, nf, 2.0, 6.0, True), mode="train",
                                    encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[5.0, 50.0],
                                    rng={k: v.cuda() for k, v in rng.items()})
    assert float(out[5].mean()) > 0.15                     # the trained field absorbs
    tg = target.cuda()
    loss = torch.nn.functional.mse_loss(out[0], tg) + torch.nn.functional.mse_loss(out[3], tg)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_o)) < 2e-3
    for m, grads in ((mc, gc), (mf, gf)):
        num = sum(float(((p.grad.cpu() - grads[k]) ** 2).sum()) for k, p in m.named_parameters())
        den = sum(float((grads[k] ** 2).sum()) for k, p in m.named_parameters())
        assert (num / den) ** 0.5 < 4e-2, (num / den) ** 0.5


# ------------------------------------------------------------------ convergence parity with the reference itself
@pytest.mark.parametrize("stochastic", [False, True])
def test_convergence_parity_with_the_reference_training_loop(tmp_path, stochastic):
    """400 iterations of the reference's own training loop (oracle/ref_train.py: the UNMODIFIED reference staged in
    oracle/_ref, torch eager fp32 on the same GPU, train_dexnerf_rgb.py:246-289) against nerf.Trainer (tcgen05 kernels,
    bf16 operands, fused Adam) from the same initial weights on the same batches of a teacher-rendered scene, with
    deterministic sampling (no jitter, no sigma noise: the only difference is the arithmetic) and with both on (each
    side draws its own): the final losses agree and the held-out view's PSNR is within 0.3 dB - the per-tensor gradient
    deviation of the bf16 contract (up to 12 % on early layers) does not change where training
    goes.  Measured: 20.25 dB vs 20.25 dB deterministic, 21.25 dB vs the reference's 20.91 dB with jitter and noise."""
    import json
    import math
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "oracle", "_ref", "nerf")):
        pytest.skip("oracle/_ref is not staged (python oracle/make_ref.py where /root/reference exists)")
    kw = dict(num_layers=4, hidden_size=128, skip_connect_every=4, num_encoding_fn_xyz=10, num_encoding_fn_dir=4)
    torch.manual_seed(21)
    teacher = nerf.FlexibleNeRFModel(**kw)
    with torch.no_grad():
        teacher.fc_alpha.weight.mul_(60.0)
        teacher.fc_alpha.bias.fill_(0.3)
        teacher.fc_rgb.weight.mul_(6.0)
    teacher = teacher.cuda()
    H = W = 24
    focal, near, far, nc, nf = 30.0, 2.0, 6.0, 32, 32
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    K = torch.tensor([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]])
    cfg_val = make_cfg(nc, nf, near, far, False, 0.0, False)

    def view(theta, phi):
        T = O.pose_spherical_world2cam(theta, phi, 4.0)
        ro, rd = O.get_ray_bundle(H, W, None, T, K)
        ro, rd = ro.reshape(-1, 3).cuda(), rd.reshape(-1, 3).cuda()
        with torch.no_grad():
            out = nerf.run_one_iter_of_nerf(H, W, focal, teacher, teacher, ro, rd, cfg_val, mode="validation",
                                            encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[])
        return ro.cpu(), rd.cpu(), out[3][..., :3].clamp(0, 1).cpu()

    views = [view(th, ph) for th, ph in ((0, -30), (60, -20), (120, -35), (180, -25), (240, -30), (300, -15))]
    ro, rd, target = (torch.cat([v[i] for v in views], 0) for i in range(3))
    val_ro, val_rd, val_target = view(30.0, -28.0)
    assert float(target.std()) > 0.05                                   # the scene has structure to learn
    iters, batch = 400, 512
    g = torch.Generator().manual_seed(9)
    batches = [torch.randint(0, ro.shape[0], (batch,), generator=g) for _ in range(iters)]
    torch.manual_seed(33)
    sc, sf = nerf.FlexibleNeRFModel(**kw), nerf.FlexibleNeRFModel(**kw)
    opts = dict(H=H, W=W, focal=focal, near=near, far=far, num_coarse=nc, num_fine=nf, perturb=stochastic,
                noise_std=0.2 if stochastic else 0.0,
                lr=5e-3, lr_decay=250, lr_decay_factor=0.1, seed=5)
    data = dict(model_kwargs=kw, init_coarse=sc.state_dict(), init_fine=sf.state_dict(), options=opts, ro=ro, rd=rd,
                target=target, batches=batches, val_ro=val_ro, val_rd=val_rd, val_target=val_target)
    path = str(tmp_path / "run.pt")
    torch.save(data, path)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "ref_train.py"), "--data", path, "--device", "cuda"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = json.loads(r.stdout.strip().splitlines()[-1])
    # this package: the same loop on nerf.Trainer
    sc, sf = sc.cuda(), sf.cuda()
    cfg = make_cfg(nc, nf, near, far, False, 0.2 if stochastic else 0.0, stochastic)
    trainer = nerf.Trainer(sc, sf, cfg, ex, ed, lr=5e-3, lr_decay=250, lr_decay_factor=0.1)
    ro_d, rd_d, tg_d = ro.cuda(), rd.cuda(), target.cuda()
    losses = []
    for idx in batches:
        idx = idx.cuda()
        losses.append(float(trainer.step(ro_d[idx], rd_d[idx], tg_d[idx])[0]))
    trainer.sync_to_modules()
    with torch.no_grad():
        out = nerf.run_one_iter_of_nerf(H, W, focal, sc, sf, val_ro.cuda(), val_rd.cuda(), cfg_val, mode="validation",
                                        encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[])
    mse = float(torch.nn.functional.mse_loss(out[3][..., :3], val_target.cuda()))
    psnr = -10.0 * math.log10(mse)
    final = sum(losses[-20:]) / 20
    print("convergence: reference final loss %.5f val PSNR %.2f dB | this package final loss %.5f val PSNR %.2f dB"
          % (ref["final_loss"], ref["val_psnr"], final, psnr))
    assert ref["losses"][0] > 4 * ref["final_loss"] and losses[0] > 4 * final        # both really trained
    assert abs(math.log(final / ref["final_loss"])) < 0.25                           # same loss level (+-25 %)
    if stochastic:      # each side draws its own jitter and noise: no worse than the reference by 0.3 dB, and close
        assert psnr > ref["val_psnr"] - 0.3 and abs(psnr - ref["val_psnr"]) < 1.0, (psnr, ref["val_psnr"])
    else:               # deterministic sampling: measured 20.25 dB against 20.25 dB, final loss 0.02705 against 0.02705
        assert abs(psnr - ref["val_psnr"]) < 0.3, (psnr, ref["val_psnr"])
