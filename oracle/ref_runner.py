"""Drive the staged UNMODIFIED reference (oracle/_ref, made by oracle/make_ref.py) on the headline workload:
`bench.py --impl reference` (CPU, all host threads) and the GPU-eager baseline (`device="cuda"`: the same
reference code as a stream of eager ATen kernels on the same B200 - the only informative competitor).

Runs in its OWN process: the reference's package is called `nerf`, like this repository's drop-in, so the two
must never share an interpreter.  Test / measurement infrastructure only.

    python oracle/ref_runner.py --device cpu|cuda --rays N --reps R [--scene c2|c3|c5] [--chunksize C]

prints one JSON line: {"rays_per_s": ..., "ms": [...], "device": ..., "cores": ..., "sha_ok": ...}."""
import argparse
import json
import math
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

SCENES = {   # bench.py's scenes
    "c2": dict(H=800, W=800, NC=64, NF=128, NEAR=2.0, FAR=6.0, FX=1111.1, MODEL=(8, 256, 4, 10, 4), ALPHA=1.0),
    "c3": dict(H=270, W=480, NC=64, NF=64, NEAR=0.3, FAR=4.0, FX=1386.4 / 4, MODEL=(8, 128, 3, 10, 4), ALPHA=1000.0),
    "c5": dict(H=720, W=1280, NC=128, NF=256, NEAR=0.3, FAR=4.0, FX=1386.4, MODEL=(8, 256, 4, 10, 4), ALPHA=1.0),
}


def load_reference():
    if not os.path.isdir(os.path.join(REF, "nerf")):
        raise SystemExit("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    sys.path.insert(0, os.path.join(REF, "_shims"))
    sys.path.insert(0, REF)
    for p in list(sys.path):          # never this repository's drop-in of the same name
        if p.rstrip("/").endswith("dex-nerf_b200"):
            sys.path.remove(p)
    import nerf as ref
    assert os.path.dirname(os.path.abspath(ref.__file__)) == os.path.join(REF, "nerf"), ref.__file__
    return ref


def repaired_flexible(ref):
    import torch

    class RepairedFlexible(ref.models.FlexibleNeRFModel):
        """models.py:233-256 with the skip condition its own __init__ builds (models.py:210) - the reference's
        forward raises AttributeError for 8-layer networks.  Parameters, names and
        init order come from the reference's __init__."""

        def forward(self, x):
            xyz, view = x[..., : self.dim_xyz], x[..., self.dim_xyz:]
            x = self.layer1(xyz)
            for i in range(len(self.layers_xyz)):
                if i % self.skip_connect_every == 0 and i > 0:
                    x = torch.cat((x, xyz), dim=-1)
                x = self.relu(self.layers_xyz[i](x))
            feat = self.relu(self.fc_feat(x))
            alpha = self.fc_alpha(x)
            x = torch.cat((feat, view), dim=-1)
            for layer in self.layers_dir:
                x = self.relu(layer(x))
            return torch.cat((self.fc_rgb(x), alpha), dim=-1)

    return RepairedFlexible


def pose_spherical_world2cam(theta, phi, radius):
    """load_blender.py:33-38 pose_spherical in the fork's OpenCV world->cam convention (same as
    oracle/nerf_oracle.py pose_spherical_world2cam; restated here so that this process imports nothing else)."""
    import torch
    t = torch.eye(4, dtype=torch.float64)
    t[2, 3] = radius
    p, th = phi / 180.0 * math.pi, theta / 180.0 * math.pi
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(p), -math.sin(p), 0], [0, math.sin(p), math.cos(p), 0], [0, 0, 0, 1]],
                      dtype=torch.float64)
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0], [0, 0, 0, 1]],
                      dtype=torch.float64)
    c2w = rt @ rp @ t
    c2w = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float64) @ c2w
    flip = torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], dtype=torch.float64))   # OpenGL camera -> OpenCV camera
    return torch.linalg.inv(c2w @ flip).to(torch.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scene", default="c2")
    ap.add_argument("--chunksize", type=int, default=4096)
    args = ap.parse_args()
    import torch
    ref = load_reference()
    sys.path.insert(0, HERE)
    import make_ref
    sc = SCENES[args.scene]
    dev = torch.device(args.device)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Flex = repaired_flexible(ref)
    torch.manual_seed(42)          # construction order of train_dexnerf_rgb.py:122-140: coarse, then fine
    L, Hd, skip, Lx, Ld = sc["MODEL"]
    mc = Flex(num_layers=L, hidden_size=Hd, skip_connect_every=skip, num_encoding_fn_xyz=Lx, num_encoding_fn_dir=Ld)
    mf = Flex(num_layers=L, hidden_size=Hd, skip_connect_every=skip, num_encoding_fn_xyz=Lx, num_encoding_fn_dir=Ld)
    if sc["ALPHA"] != 1.0:
        with torch.no_grad():
            for m in (mc, mf):
                m.fc_alpha.weight.mul_(sc["ALPHA"])
                m.fc_alpha.bias.mul_(sc["ALPHA"])
    mc, mf = mc.to(dev), mf.to(dev)
    H, W = sc["H"], sc["W"]
    T = pose_spherical_world2cam(30.0, -30.0, 4.0).to(dev)
    K = torch.tensor([[sc["FX"], 0.0, W / 2.0], [0.0, sc["FX"], H / 2.0], [0.0, 0.0, 1.0]]).to(dev)
    mode = dict(chunksize=args.chunksize, perturb=False, num_coarse=sc["NC"], num_fine=sc["NF"], white_background=False,
                radiance_field_noise_std=0.0, lindisp=False)
    cfg = ref.CfgNode(dict(dataset=dict(no_ndc=True, near=sc["NEAR"], far=sc["FAR"]),
                           nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
    ex, ed = ref.get_embedding_function(Lx, True, True), ref.get_embedding_function(Ld, True, True)
    thr = [float(m) for m in range(5, 105, 5)]
    rows = max(1, min(H, (args.rays + W - 1) // W))
    r0 = max(0, H // 2 - rows // 2)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    def once():
        with torch.no_grad():
            ro, rd = ref.get_ray_bundle(H, W, None, T, K)            # the reference makes the whole bundle
            ro, rd = ro[r0:r0 + rows], rd[r0:r0 + rows]
            out = ref.run_one_iter_of_nerf(H, W, sc["FX"], mc, mf, ro, rd, cfg, mode="validation",
                                           encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=thr)
        sync()
        return out

    for _ in range(args.warmup):
        once()
    ms = []
    for _ in range(args.reps):
        sync()
        t0 = time.perf_counter()
        out = once()
        ms.append(1e3 * (time.perf_counter() - t0))
    n = rows * W
    print(json.dumps({"rays_per_s": n / (statistics.median(ms) * 1e-3), "ms": ms, "rays": n, "device": str(dev),
                      "cores": cores, "chunksize": args.chunksize, "scene": args.scene, "sha_ok": make_ref.verify(),
                      "outputs": len(out), "acc_fine_mean": float(out[5].mean())}))


if __name__ == "__main__":
    main()
