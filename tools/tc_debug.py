"""Bring-up aid for the tcgen05 MLP kernel: dumps the raw fp32 accumulator of every (layer, pass)
through the kernel's debug tap and prints its error against a bf16-operand emulation in torch.
Run on the GPU box:  python tools/tc_debug.py [hidden] [n_rays] [S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))

import nerf  # noqa: E402
from nerf import tensorcore  # noqa: E402
from oracle import nerf_oracle as O  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def emulate(model, enc_xyz, enc_dir):
    """Per tensor-core layer: pre-bias accumulator; plus the final (rgb, sigma)."""
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    skip = model.skip_connect_every
    n_trunk = len(model.layers_xyz)
    pre = []
    xyz, dr = bf(enc_xyz), bf(enc_dir)
    a = xyz @ bf(sd["layer1.weight"]).t()
    pre.append(a)
    h = a + sd["layer1.bias"]
    for i in range(n_trunk):
        inp = bf(h)
        if i % skip == 0 and i > 0:
            inp = torch.cat((inp, xyz), -1)
        a = inp @ bf(sd[f"layers_xyz.{i}.weight"]).t()
        pre.append(a)
        h = torch.relu(a + sd[f"layers_xyz.{i}.bias"])
    sigma = h @ sd["fc_alpha.weight"].t() + sd["fc_alpha.bias"]
    a = bf(h) @ bf(sd["fc_feat.weight"]).t()
    pre.append(a)
    feat = torch.relu(a + sd["fc_feat.bias"])
    a = torch.cat((bf(feat), dr), -1) @ bf(sd["layers_dir.0.weight"]).t()
    pre.append(a)
    y = torch.relu(a + sd["layers_dir.0.bias"])
    rgb = y @ sd["fc_rgb.weight"].t() + sd["fc_rgb.bias"]
    return pre, torch.cat((rgb, sigma), -1)


def main():
    hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    skip = 4 if hidden == 256 else 3
    torch.manual_seed(1)
    model = nerf.FlexibleNeRFModel(8, hidden, skip, 10, 4)
    with torch.no_grad():
        model.fc_alpha.weight.mul_(50.0)
    model = model.cuda()
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    prog = model.program(ex, ed)
    g = torch.Generator().manual_seed(0)
    ro = torch.randn(n, 3, generator=g) * 0.3
    rd = torch.randn(n, 3, generator=g)
    vd = rd / rd.norm(dim=-1, keepdim=True)
    z = torch.sort(2 + 4 * torch.rand(n, S, generator=g), dim=-1).values
    pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3)
    enc_xyz = O.positional_encoding(pts, 10)
    enc_dir = O.positional_encoding(vd[:, None, :].expand(n, S, 3).reshape(-1, 3), 4)
    pre, out = emulate(model, enc_xyz, enc_dir)
    M = n * S
    Mp = (M + 255) // 256 * 256
    n_layers = len(pre)
    print("hidden", hidden, "rays", n, "S", S, "samples", M, "tc layers", n_layers, flush=True)
    rf = torch.empty(n, S, 4, device="cuda")
    worst = 0.0
    for l in range(n_layers):
        width = pre[l].shape[1]
        for p in range((width + 127) // 128):
            dbg = torch.full((Mp, 128), float("nan"), device="cuda")
            tensorcore.query(model, prog, ro.cuda(), rd.cuda(), vd.cuda(), z.cuda(), rf, dbg=dbg, dbg_layer=l, dbg_pass=p)
            torch.cuda.synchronize()
            cols = min(128, width - p * 128)
            got = dbg[:M, :cols].cpu()
            ref = pre[l][:, p * 128:p * 128 + cols]
            err = (got - ref).abs()
            scale = float(ref.abs().max())
            nan = int(torch.isnan(got).sum())
            print("layer %2d pass %d: max|err| %.3e (ref max %.3e) rel %.2e nan %d | row-wise worst row %d col %d"
                  % (l, p, float(err.nan_to_num(1e9).max()), scale, float(err.nan_to_num(1e9).max()) / max(scale, 1e-9),
                     nan, int(err.nan_to_num(1e9).max(1).values.argmax()), int(err.nan_to_num(1e9).max(0).values.argmax())),
                  flush=True)
            if l == 0 and p == 0:
                print("   got[0,:6]", got[0, :6].tolist(), "\n   ref[0,:6]", ref[0, :6].tolist(), flush=True)
                print("   got[1,:6]", got[1, :6].tolist(), "\n   ref[1,:6]", ref[1, :6].tolist(), flush=True)
            worst = max(worst, float(err.nan_to_num(1e9).max()) / max(scale, 1e-9))
    err = (rf.reshape(-1, 4).cpu() - out).abs()
    print("final rf: max|err| rgb %.3e sigma %.3e (ref max %.3e / %.3e)"
          % (float(err[:, :3].max()), float(err[:, 3].max()), float(out[:, :3].abs().max()), float(out[:, 3].abs().max())))
    fp32 = O.flexible_forward({k: v.detach().float().cpu() for k, v in model.state_dict().items()},
                              torch.cat((enc_xyz, enc_dir), -1), skip_connect_every=skip)
    e32 = (rf.reshape(-1, 4).cpu() - fp32).abs()
    print("vs fp32 oracle: rgb %.3e sigma %.3e" % (float(e32[:, :3].max()), float(e32[:, 3].max())))
    print("WORST_REL", worst)


if __name__ == "__main__":
    main()
