"""Train the staged UNMODIFIED reference (oracle/_ref) with the statements of its own training loop
(train_dexnerf_rgb.py:246-289: run_one_iter_of_nerf(mode="train"), mse(coarse) + mse(fine), loss.backward(),
torch.optim.Adam step, exponential learning-rate decay) on a data file written by the convergence-parity test, from
given initial weights and a given sequence of ray batches; prints one JSON line with the loss curve and the PSNR of a
held-out view.  Own process (the reference's package is called `nerf` like this repository's drop-in).  Test
infrastructure only.

    python oracle/ref_train.py --data run.pt --device cuda|cpu"""
import argparse
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_runner  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", required=True)
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    import torch
    ref = ref_runner.load_reference()
    Flex = ref_runner.repaired_flexible(ref)
    d = torch.load(args.data, map_location="cpu")
    dev = torch.device(args.device)
    kw = d["model_kwargs"]
    mc, mf = Flex(**kw), Flex(**kw)
    mc.load_state_dict(d["init_coarse"]); mf.load_state_dict(d["init_fine"])
    mc, mf = mc.to(dev), mf.to(dev)
    o = d["options"]
    mode = dict(chunksize=1 << 20, perturb=o["perturb"], num_coarse=o["num_coarse"], num_fine=o["num_fine"],
                white_background=False, radiance_field_noise_std=o["noise_std"], lindisp=False)
    val = dict(mode, perturb=False, radiance_field_noise_std=0.0)
    cfg = ref.CfgNode(dict(dataset=dict(no_ndc=True, near=o["near"], far=o["far"]),
                           nerf=dict(use_viewdirs=True, train=mode, validation=val)))
    ex = ref.get_embedding_function(kw["num_encoding_fn_xyz"], True, True)
    ed = ref.get_embedding_function(kw["num_encoding_fn_dir"], True, True)
    ro, rd, target = d["ro"].to(dev), d["rd"].to(dev), d["target"].to(dev)
    opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=o["lr"])     # train_dexnerf_rgb.py:142-148
    torch.manual_seed(o["seed"])
    losses = []
    for i, idx in enumerate(d["batches"]):
        idx = idx.to(dev)
        out = ref.run_one_iter_of_nerf(o["H"], o["W"], o["focal"], mc, mf, ro[idx], rd[idx], cfg, mode="train",
                                       encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[])
        loss = torch.nn.functional.mse_loss(out[0][..., :3], target[idx]) + \
            torch.nn.functional.mse_loss(out[3][..., :3], target[idx])
        loss.backward()
        opt.step()
        opt.zero_grad()
        lr_new = o["lr"] * (o["lr_decay_factor"] ** (i / (o["lr_decay"] * 1000)))          # :283-289
        for g in opt.param_groups:
            g["lr"] = lr_new
        losses.append(float(loss))
    with torch.no_grad():
        out = ref.run_one_iter_of_nerf(o["H"], o["W"], o["focal"], mc, mf, d["val_ro"].to(dev), d["val_rd"].to(dev), cfg,
                                       mode="validation", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=[])
        mse = float(torch.nn.functional.mse_loss(out[3][..., :3], d["val_target"].to(dev)))
    import math
    print(json.dumps({"losses": losses[::10], "final_loss": sum(losses[-20:]) / 20, "val_mse": mse,
                      "val_psnr": -10.0 * math.log10(mse), "iters": len(losses), "device": str(dev)}))


if __name__ == "__main__":
    main()
