#!/usr/bin/env python
"""Headline benchmark: rays/s rendered by the full-paper NeRF (BASELINE.json config 2):
800x800 frame, 64 coarse + 128 fine samples, two 8x256 skip-4 FlexibleNeRFModels (L=10/4, view
directions), T=20 Dex-NeRF thresholds, validation mode, random-init weights, synthetic camera.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" renders one frame.  With N GPUs the image rows are block-partitioned over the ranks
(no collective on the data path; total work fixed -> "scaling": "strong").  The timed region is
bracketed by barrier + synchronize; the JSON line carries the device-timed throughput (`value`),
the end-to-end throughput through the public `nerf` API with host buffers (`e2e`), the roofline
of the dominant kernel (the fine-pass MLP query) and the CPU oracle timed on this box.
`--impl reference` times the UNMODIFIED reference on the host cores (oracle/_ref, staged by
oracle/make_ref.py and shipped to the GPU box; the oracle port only if that copy is absent) on a
bounded sample of the same workload.
"""
import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))

import torch  # noqa: E402

H = W = 800
NC, NF = 64, 128
NEAR, FAR = 2.0, 6.0
FX = 1111.1
THRESHOLDS = [float(m) for m in range(5, 105, 5)]
FLOP_PER_EVAL = 2 * 593408                        # SURVEY.md section 8(d), unpadded
FLOP_PER_RAY = (NC + NC + NF) * FLOP_PER_EVAL     # 303 824 896
CPU_SAMPLE_RAYS = 4096
MODEL_ARGS = (8, 256, 4, 10, 4)                   # FlexibleNeRFModel(num_layers, hidden, skip, Lx, Ld)
METRIC = "rays/sec render (64+128 samples, 8x256 MLP)"
WORKLOAD = ("C2: full-paper NeRF render 800x800, 64 coarse + 128 fine samples, two 8x256 skip-4 "
            "FlexibleNeRFModels (L=10/4, viewdirs), T=20 Dex thresholds, validation mode, random-init")
SCENE = "c2"

# Secondary render scenes (parity-test configurations of BASELINE.json, measured
# with the same step and the same JSON line so that their HBM-side kernels are on record too):
#   c3  Dex-NeRF sigma-threshold depth render at the messytable size: 270x480, 64+64 samples, 8x128 skip-3
#       (config/messytable-obj.yml:36-53), near 0.3 / far 4, T=20
#   c5  IR variant at 1280x720 with 128+256 samples (train_nerf_ir.py; config/messytable-obj-edward.yml),
#       8x256 skip-4, near 0.3 / far 4, T=20 - the configuration that stresses sample_pdf and compositing
C3_ALPHA_SCALE = 1000.0
SCENES = {
    "c3": dict(H=270, W=480, NC=64, NF=64, NEAR=0.3, FAR=4.0, FX=1386.4 / 4, MODEL_ARGS=(8, 128, 3, 10, 4),
               METRIC="rays/sec render (64+64 samples, 8x128 MLP, Dex depth)",
               WORKLOAD="C3: Dex-NeRF sigma-threshold depth render 270x480, 64+64 samples, two 8x128 skip-3 "
                        "FlexibleNeRFModels (L=10/4, viewdirs), T=20 Dex thresholds, near 0.3 / far 4, "
                        "validation mode, random-init with fc_alpha x1000 (sigma spans the thresholds)"),
    "c5": dict(H=720, W=1280, NC=128, NF=256, NEAR=0.3, FAR=4.0, FX=1386.4, MODEL_ARGS=(8, 256, 4, 10, 4),
               METRIC="rays/sec render (128+256 samples, 8x256 MLP)",
               WORKLOAD="C5: Dex-NeRF IR variant render 1280x720, 128 coarse + 256 fine samples, two 8x256 skip-4 "
                        "FlexibleNeRFModels (L=10/4, viewdirs), T=20 Dex thresholds, near 0.3 / far 4, "
                        "validation mode, random-init"),
}


def apply_scene(name):
    """Point the module-level workload constants at a secondary scene."""
    global H, W, NC, NF, NEAR, FAR, FX, MODEL_ARGS, METRIC, WORKLOAD, SCENE, FLOP_PER_EVAL, FLOP_PER_RAY
    sc = SCENES[name]
    H, W, NC, NF, NEAR, FAR, FX = sc["H"], sc["W"], sc["NC"], sc["NF"], sc["NEAR"], sc["FAR"], sc["FX"]
    MODEL_ARGS, METRIC, WORKLOAD, SCENE = sc["MODEL_ARGS"], sc["METRIC"], sc["WORKLOAD"], name
    mc, _ = state_dicts()
    FLOP_PER_EVAL = 2 * sum(p.numel() for k, p in mc.named_parameters() if k.endswith("weight"))
    FLOP_PER_RAY = (NC + NC + NF) * FLOP_PER_EVAL


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def ncu_traffic(key, rays):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu
    capture (profiles/ncu_traffic.json), scaled to this launch's ray count; None when absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get(key)
    if not d:
        return None
    return (d["dram_bytes_read"] + d["dram_bytes_write"]) * rays / d["rays"]


def train_traffic():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p)).get("train_fine_c4")
    return None if not d else d["mlp_tc_train_fwd"] + d["mlp_tc_bwd_dx"] + d["mlp_tc_bwd_dw"]


def camera():
    """pose_spherical(30, -30, 4) of load_blender.py:33-38 in the fork's
    world->cam convention (host-side helpers of the product package; the oracle is only touched by
    the CPU legs below)."""
    import nerf
    T = nerf.world2cam_from_blender_pose(nerf.pose_spherical(30.0, -30.0, 4.0))
    K = torch.tensor([[FX, 0.0, W / 2.0], [0.0, FX, H / 2.0], [0.0, 0.0, 1.0]])
    return T, K


def make_cfg(nerf):
    mode = dict(chunksize=1 << 30, perturb=False, num_coarse=NC, num_fine=NF, white_background=False,
                radiance_field_noise_std=0.0, lindisp=False)
    return nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=NEAR, far=FAR),
                             nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))


def state_dicts():
    """Random-init weights of the C2 pair, drawn on the CPU under torch.manual_seed(42) exactly as
    the reference scripts would construct them (coarse first, then fine)."""
    import nerf
    torch.manual_seed(42)
    mc = nerf.FlexibleNeRFModel(*MODEL_ARGS)
    mf = nerf.FlexibleNeRFModel(*MODEL_ARGS)
    if SCENE == "c3":
        # a random-init field is nearly empty; scale fc_alpha so that sigma
        # spans the 5...100 threshold range and the first-crossing logic has crossings to find
        with torch.no_grad():
            for m in (mc, mf):
                m.fc_alpha.weight.mul_(C3_ALPHA_SCALE)
                m.fc_alpha.bias.mul_(C3_ALPHA_SCALE)
    return mc, mf


class ClockSampler:
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=100):
        self.proc, self.idx, self.period_ms = None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self, window=None):
        """window = (t0, t1) in time.time() seconds: keep only the samples taken inside it (the sampler may have been
        started earlier so that its start-up is not inside the timed region)."""
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if window is not None:
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    continue
                if ts < window[0] - 0.05 or ts > window[1] + 0.05:
                    continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    samples=len(sm), reasons=sorted(reasons))


def cpu_oracle_rays_per_s(n_rays=CPU_SAMPLE_RAYS, reps=1):
    """Fallback CPU leg when oracle/_ref is absent: the oracle (CPU port of the reference's algorithm) on a bounded
    sample of the workload: `n_rays` rays from the centre rows of the camera, all host threads."""
    from oracle import nerf_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    T, K = camera()
    ro, rd = O.get_ray_bundle(H, W, None, T, K)
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    start = (H // 2) * W
    mc, mf = state_dicts()
    sdc = {k: v.detach() for k, v in mc.state_dict().items()}
    sdf = {k: v.detach() for k, v in mf.state_dict().items()}
    opts = O.RenderOptions(near=NEAR, far=FAR, num_coarse=NC, num_fine=NF, Lx=10, Ld=4, chunksize=REF_CHUNK)
    fc = lambda x: O.flexible_forward(sdc, x, skip_connect_every=MODEL_ARGS[2])   # noqa: E731
    ff = lambda x: O.flexible_forward(sdf, x, skip_connect_every=MODEL_ARGS[2])   # noqa: E731
    with torch.no_grad():
        O.render_rays(ro[start:start + 128], rd[start:start + 128], fc, ff, opts, THRESHOLDS)   # warm-up
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            O.render_rays(ro[start:start + n_rays], rd[start:start + n_rays], fc, ff, opts, THRESHOLDS)
            times.append(time.perf_counter() - t0)
    return n_rays / statistics.median(times), torch.get_num_threads(), [1e3 * x for x in times]


REF_RAYS = 4096        # >= 4096 rays x 3 repetitions after one warm-up, chunksize 4096
REF_CHUNK = 4096
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def run_ref_runner(device, rays, reps, warmup, chunksize):
    """The staged UNMODIFIED reference (oracle/_ref, recipe oracle/make_ref.py) in its own process - its package is
    called `nerf` like this repository's.  Returns the runner's JSON dict or None."""
    cmd = [sys.executable, os.path.join(ROOT, "oracle", "ref_runner.py"), "--device", device, "--rays", str(rays),
           "--reps", str(reps), "--warmup", str(warmup), "--scene", SCENE, "--chunksize", str(chunksize)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
        if r.returncode != 0:
            sys.stderr.write("ref_runner failed: %s\n" % r.stderr[-2000:])
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:          # noqa: BLE001
        sys.stderr.write("ref_runner failed: %r\n" % (e,))
        return None


def cpu_baseline(reps=3, warmup=1):
    """The reference's CPU path on this box's host cores on a bounded sample (REF_RAYS rays of the centre rows, the
    full coarse -> fine pipeline): the staged reference itself when oracle/_ref travelled here ("reference"), else
    the oracle port ("port")."""
    d = run_ref_runner("cpu", REF_RAYS, reps, warmup, REF_CHUNK) if reference_available() else None
    if d is not None:
        return {"value": d["rays_per_s"], "unit": "rays/s", "cores": d["cores"], "kind": "reference",
                "sample": "%d rays of the same %s frame (centre rows), chunksize %d, median of %d repetitions after %d "
                          "warm-up; the unmodified reference (oracle/_ref, sha256 manifest %s) with the repaired 8-layer "
                          "forward" % (d["rays"], SCENE.upper(), REF_CHUNK, reps, warmup, "ok" if d["sha_ok"] else "MISMATCH"),
                "ms": d["ms"]}
    v, cores, ms = cpu_oracle_rays_per_s(REF_RAYS, reps=reps)
    return {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": "%d rays of the same %s frame (centre rows), chunksize %d, median of %d repetitions after a 128-ray "
                      "warm-up; oracle port (oracle/_ref not present)" % (REF_RAYS, SCENE.upper(), REF_CHUNK, reps), "ms": ms}


def gpu_eager_baseline():
    """The same unmodified reference code with device="cuda": a stream of eager ATen kernels on this B200 (the
    reference has no GPU-specific code of its own).  ~A sixth of a frame per repetition, the YAMLs' chunksize."""
    if not reference_available():
        return None
    d = run_ref_runner("cuda", 131072, 3, 1, 131072)
    if d is None:
        return None
    return {"value": d["rays_per_s"], "unit": "rays/s", "kind": "reference code, torch eager on cuda:0",
            "sample": "%d rays of the same %s frame (centre rows), chunksize 131072 (config/lego.yml:136), median of 3 "
                      "repetitions after 1 warm-up" % (d["rays"], SCENE.upper()), "ms": d["ms"]}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, each step a bounded
    sample (REF_RAYS rays) of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_baseline(reps=args.steps, warmup=args.warmup)
    ms = sum(base["ms"]) / len(base["ms"])                                   # mean step, like the GPU arm
    rays = int(round(base["value"] * statistics.median(base["ms"]) * 1e-3))  # rays per step (whole image rows)
    value = rays / (ms / 1e3)
    base = dict(base, value=value)
    base.pop("ms", None)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


TRAIN_RAYS = 4096
TRAIN_FLOP_PER_RAY = 3 * FLOP_PER_RAY          # forward + dX + dW: 911.5 MFLOP/ray


def run_train(args, dist, rank, world, dev, quiet=False):
    """BASELINE config 4: one training iteration per step on 4096 rays per GPU drawn from the C2
    bundle (np.random.seed(42)), train mode (perturb, noise 0.2, random u), loss = mse(coarse) +
    mse(fine), backward through the tensor-core kernels, one flat NCCL all-reduce of both MLPs'
    gradients (N > 1), Adam(lr 5e-3) with the reference's exponential decay.  Weak scaling."""
    import numpy as np
    import nerf
    from nerf import _lib as L
    from nerf import training as TR
    mc, mf = state_dicts()
    mc, mf = mc.to(dev), mf.to(dev)
    mode = dict(chunksize=1 << 30, perturb=True, num_coarse=NC, num_fine=NF, white_background=False,
                radiance_field_noise_std=0.2, lindisp=False)
    cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=NEAR, far=FAR),
                            nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    T_cpu, K_cpu = camera()
    ro, rd = nerf.get_ray_bundle(H, W, None, T_cpu.to(dev), K_cpu.to(dev))
    ro, rd = ro.reshape(-1, 3), rd.reshape(-1, 3)
    np.random.seed(42 + rank)
    g = torch.Generator().manual_seed(1234 + rank)
    n_batches = 4
    sel = [torch.from_numpy(np.random.choice(H * W, size=TRAIN_RAYS, replace=False)) for _ in range(n_batches)]
    tgt_host = [torch.rand(TRAIN_RAYS, 3, generator=g).pin_memory() for _ in range(n_batches)]
    sel_host = [s.pin_memory() for s in sel]
    sel_dev = [s.to(dev) for s in sel]
    tgt_dev = [x.to(dev) for x in tgt_host]
    api = getattr(args, "train_api", "trainer")
    if api == "trainer":
        # nerf.Trainer: flat parameter / gradient / Adam buffers, one fused Adam launch, one flat all-reduce
        trainer = nerf.Trainer(mc, mf, cfg, ex, ed, lr=5e-3, lr_decay=250, lr_decay_factor=0.1, world_size=world)

        def step(sel_i, tgt_i):
            return trainer.step(ro[sel_i], rd[sel_i], tgt_i)[0]
    else:
        # the reference's loop verbatim: autograd through run_one_iter_of_nerf + torch.optim.Adam
        opt = torch.optim.Adam(list(mc.parameters()) + list(mf.parameters()), lr=5e-3)
        it = [0]

        def step(sel_i, tgt_i):
            for pg in opt.param_groups:
                pg["lr"] = nerf.learning_rate(5e-3, it[0], 250, 0.1)
            it[0] += 1
            return nerf.train_step(mc, mf, opt, ro[sel_i], rd[sel_i], tgt_i, cfg, ex, ed, m_thres_cand=[],
                                   height=H, width=W, focal=FX, world_size=world)[0]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        if dist is None:
            return ms
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # The clock sampler (an nvidia-smi process) starts BEFORE the warm-up and gets 0.4 s to initialise: forking it and
    # its NVML start-up inside a ~100 ms timed region of 4 ms steps would cost the launching thread more than a step.
    sampler = ClockSampler(dev.index or 0, period_ms=int(os.environ.get("DEXNERF_BENCH_SMI_MS", "25")))
    if rank == 0 and sampler.period_ms > 0:
        sampler.start()
    time.sleep(0.4)
    for i in range(max(args.warmup, 3)):
        step(sel_dev[i % n_batches], tgt_dev[i % n_batches])
    barrier()
    from nerf import render as RD
    launches0 = L.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step(sel_dev[i % n_batches], tgt_dev[i % n_batches])
    e1.record()
    barrier()
    t_wall1 = time.time()
    launches = L.launch_count - launches0
    clocks = sampler.stop((t_wall0, t_wall1)) if rank == 0 else None
    ms_dev = reduce_max(e0.elapsed_time(e1) / args.steps)
    # per-kernel times: events recorded by the library around every launch - in steps of their own AFTER the timed
    # region (the ~40 event records per step cost launch gaps, they do not belong in `value`)
    TR.event_log = []
    if api == "trainer":
        trainer.timing = (RD.Events(), RD.Events())
    for i in range(2):
        step(sel_dev[i % n_batches], tgt_dev[i % n_batches])
    torch.cuda.synchronize(dev)
    log, TR.event_log = TR.event_log, None
    parts = {}
    for name, a, b, n, S in log:
        parts.setdefault("%s_S%d" % (name, S), []).append(a.elapsed_time(b))
    parts = {k: sum(v) / len(v) for k, v in parts.items()}
    small = {}
    if api == "trainer":
        fw, bw = trainer.timing[0].elapsed_ms(), trainer.timing[1].elapsed_ms(RD.BWD_LAUNCH_NAMES)
        trainer.timing = None
        parts = {"mlp_tc_train_fwd_S%d" % NC: fw["mlp_coarse"], "mlp_tc_train_fwd_S%d" % (NC + NF): fw["mlp_fine"]}
        if "mlp_bwd_dw_fine" in bw:      # the two backward kernels (default)
            parts.update({"mlp_tc_bwd_dx_S%d" % NC: bw["mlp_bwd_coarse"], "mlp_tc_bwd_dx_S%d" % (NC + NF): bw["mlp_bwd_fine"],
                          "mlp_tc_bwd_dw_S%d" % NC: bw["mlp_bwd_dw_coarse"], "mlp_tc_bwd_dw_S%d" % (NC + NF): bw["mlp_bwd_dw_fine"]})
        else:                            # DEXNERF_BWD=fused / shared: ONE backward launch (chain + weight-gradient GEMM)
            parts.update({"mlp_tc_bwd_fused_S%d" % NC: bw["mlp_bwd_coarse"], "mlp_tc_bwd_fused_S%d" % (NC + NF): bw["mlp_bwd_fine"]})
        small = {k: v for k, v in list(fw.items()) + list(bw.items()) if not k.startswith("mlp_")}
    # multi-GPU correctness on record: after the same number of steps every rank must hold the same parameters
    param_check = None
    if api == "trainer":
        cs = torch.stack((trainer.params.double().sum(), (trainer.params.double() ** 2).sum(),
                          trainer.params.view(torch.int32).to(torch.int64).sum().double()))
        if dist is not None:
            allcs = [torch.empty_like(cs) for _ in range(world)]
            dist.all_gather(allcs, cs)
            same = all(torch.equal(allcs[0], c) for c in allcs)
            param_check = {"ranks": world, "identical_across_ranks": bool(same), "sum": float(cs[0]), "sumsq": float(cs[1]),
                           "int_view_sum": float(cs[2])}
            assert same, "data-parallel ranks diverged: parameter checksums differ"
        else:
            param_check = {"ranks": 1, "identical_across_ranks": True, "sum": float(cs[0]), "sumsq": float(cs[1]),
                           "int_view_sum": float(cs[2])}
    # end to end: ray indices + targets come from pinned host memory, the loss goes back to the host
    loss_pin = torch.empty((), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        loss = step(sel_host[i % n_batches].to(dev, non_blocking=True), tgt_host[i % n_batches].to(dev, non_blocking=True))
        loss_pin.copy_(loss, non_blocking=True)
    for i in range(2):
        e2e_step(i)
    barrier()
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e1.record()
    barrier()
    ms_e2e = reduce_max(e0.elapsed_time(e1) / args.steps)
    if rank != 0:
        return None
    pk = peaks()
    rays = TRAIN_RAYS * world
    dw_ms = parts.get("mlp_tc_bwd_dw_S%d" % (NC + NF))
    dx_ms = parts.get("mlp_tc_bwd_dx_S%d" % (NC + NF))
    fw_ms = parts.get("mlp_tc_train_fwd_S%d" % (NC + NF))
    flop_fine = TRAIN_RAYS * (NC + NF) * FLOP_PER_EVAL
    mlp_ms = sum(v for k, v in parts.items() if k.startswith("mlp_tc"))
    line = {
        "metric": "training rays/sec (4096-ray batches per GPU, 64+128 samples, 8x256 MLP, fwd+bwd+allreduce+Adam)",
        "value": rays / (ms_dev * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C4: NeRF training iteration, 4096 rays per GPU drawn from the C2 800x800 bundle, train "
                               "mode (perturb, sigma noise 0.2), two 8x256 skip-4 FlexibleNeRFModels, mse(coarse)+mse(fine), "
                               "Adam lr 5e-3 with exponential decay; N > 1: gradient mean + Adam in one kernel over NVLink peer "
                               "memory (NCCL all-reduce with DEXNERF_P2P=0 or the autograd API)",
                   "rays_per_step": rays, "parallelism": "dp%d" % world,
                   "gradient_exchange": ("none (1 GPU)" if world == 1 else
                                         ("fused P2P all-reduce + Adam kernel (csrc/p2p.cu)" if api == "trainer" and
                                          getattr(trainer, "_p2p", None) is not None else "NCCL all-reduce")),
                   "api": "nerf.Trainer (flat buffers, fused Adam)" if api == "trainer"
                          else "run_one_iter_of_nerf(mode='train') autograd + torch.optim.Adam",
                   "l2": "the per-step forward tape (5.3 KB/sample, 5.6 GB per step) exceeds the 126 MB L2"},
        "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": TRAIN_RAYS * (8 + 12), "d2h_bytes_per_step": 4},
        "gpu_launches": launches,
        # The training step is bound by the TAPE traffic: per 128-sample tile the
        # forward writes 666 KB, the activation-gradient chain writes 612 KB (+ 34 KB of masks read) and the
        # weight-gradient GEMM reads 1 358 KB (1 424 KB before fc_alpha's gradient moved into the feature layer's item)
        # - 2 670 KB of algorithmic HBM bytes against 0.46 GFLOP.
        "roofline": {"bound": "hbm", "kernel": "mlp_tc train fwd + bwd_dx + bwd_dw (fine pass, 192 samples/ray)",
                     "achieved": None, "peak": pk["hbm"], "unit": "GB/s", "peak_source": pk["source"] + " HBM copy",
                     "traffic": train_traffic(), "kernel_ms": {k: round(v, 4) for k, v in sorted(parts.items())},
                     "mlp_share_of_step": mlp_ms / ms_dev, "flop_per_launch": flop_fine,
                     "tensor_tflops": 3 * flop_fine / ((fw_ms + dx_ms + dw_ms) * 1e-3) / 1e12 if (fw_ms and dx_ms and dw_ms) else None,
                     "tensor_frac_of_sustained_peak": None},
        "clocks": clocks,
        "param_checksum": param_check,
        "non_mlp_ms": ms_dev - mlp_ms,
        "small_kernel_ms": {k: round(v, 4) for k, v in sorted(small.items())},
    }
    bw_ms = parts.get("mlp_tc_bwd_fused_S%d" % (NC + NF))
    if fw_ms and bw_ms:
        # Fused backward: the gradient images go from the chain CTAs to the weight-gradient
        # CTAs through L2, so the step's HBM traffic is the forward tape only (written once: 666 KB per 128-sample
        # tile; read once by the GEMM, the two head operands twice: ~766 KB) and the bound is the tensor pipe:
        # forward + activation-gradient chain + weight-gradient GEMM = 1 + 0.896 + 1 forward-equivalents.
        f_dx = TRAIN_RAYS * (NC + NF) * 2.0 * (128 * 256 + 8 * 256 * 256)
        flop3 = 2 * flop_fine + f_dx
        r = line["roofline"]
        r.update({"bound": "tensor", "kernel": "mlp_tc train fwd + fused backward (fine pass, 192 samples/ray)",
                  "achieved": flop3 / ((fw_ms + bw_ms) * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                  "peak_source": pk["source"] + " bf16 sustained", "flop_per_launch": flop3})
        r["frac"] = r["achieved"] / pk["tf_sustained"]
        r["tensor_tflops"] = r["achieved"]
        r["tensor_frac_of_sustained_peak"] = r["frac"]
        tiles_ = TRAIN_RAYS * (NC + NF) / 128.0
        r["hbm_bytes_algorithmic"] = tiles_ * (666.0 + 766.0) * 1024.0
        r["hbm_gbs"] = r["hbm_bytes_algorithmic"] / ((fw_ms + bw_ms) * 1e-3) / 1e9
        r["hbm_frac_of_peak"] = r["hbm_gbs"] / pk["hbm"]
        r["kernels"] = {
            "mlp_tc_train_fwd": {"ms": fw_ms, "tflops": flop_fine / (fw_ms * 1e-3) / 1e12,
                                 "hbm_write_gbs": tiles_ * 666.0 * 1024.0 / (fw_ms * 1e-3) / 1e9},
            "mlp_tc_bwd_fused": {"ms": bw_ms, "tflops": (flop_fine + f_dx) / (bw_ms * 1e-3) / 1e12,
                                 "tensor_frac_of_sustained_peak": (flop_fine + f_dx) / (bw_ms * 1e-3) / 1e12 / pk["tf_sustained"],
                                 "hbm_read_gbs": tiles_ * 766.0 * 1024.0 / (bw_ms * 1e-3) / 1e9}}
        if not quiet:
            print(json.dumps(line))
        return line
    if fw_ms and dx_ms and dw_ms:
        tape_bytes = TRAIN_RAYS * (NC + NF) / 128.0 * (666.0 + 646.0 + 1358.0) * 1024.0
        r = line["roofline"]
        r["bytes_per_launch"] = tape_bytes
        r["achieved"] = tape_bytes / ((fw_ms + dx_ms + dw_ms) * 1e-3) / 1e9
        r["frac"] = r["achieved"] / pk["hbm"]
        r["tensor_frac_of_sustained_peak"] = r["tensor_tflops"] / pk["tf_sustained"]
    # per kernel: tensor rate AND tape bandwidth (forward / activation-gradient chain write the tape while they
    # compute; the weight-gradient GEMM only streams it); bytes = algorithmic tape bytes per launch
    tiles = TRAIN_RAYS * (NC + NF) / 128.0
    per = {}

    def entry(ms, flop, kbytes_per_tile, bound):
        b = tiles * kbytes_per_tile * 1024.0
        return {"bound": bound, "ms": ms, "tflops": flop / (ms * 1e-3) / 1e12,
                "tensor_frac_of_sustained_peak": flop / (ms * 1e-3) / 1e12 / pk["tf_sustained"],
                "gbs": b / (ms * 1e-3) / 1e9, "hbm_frac_of_peak": b / (ms * 1e-3) / 1e9 / pk["hbm"], "bytes": b}
    if fw_ms:
        per["mlp_tc_train_fwd"] = entry(fw_ms, flop_fine, 666.0, "crossbar write port (32 B/clk/SM) + tensor")
    if dx_ms:
        f = TRAIN_RAYS * (NC + NF) * 2.0 * (128 * 256 + 8 * 256 * 256)      # dir^T(feat part) + fc_feat^T + 7 trunk^T
        per["mlp_tc_bwd_dx"] = entry(dx_ms, f, 646.0, "crossbar write port (32 B/clk/SM) + tensor")
    if dw_ms:
        per["mlp_tc_bwd_dw"] = entry(dw_ms, flop_fine, 1358.0, "hbm read")
    line["roofline"]["kernels"] = per
    if not quiet:
        print(json.dumps(line))
    return line


def workload_config(n_gpus):
    return {"workload": WORKLOAD,
            "rays_per_step": H * W, "rows_per_gpu": H // n_gpus, "parallelism": "rows%d" % n_gpus,
            "l2": "per-step intermediates (z_fine + radiance field, >2 GB/GPU at N=1) exceed the 126 MB L2; "
                  "a 256 MB buffer is also rewritten between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=None, choices=[None, "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the short C4 training measurement of the default run")
    ap.add_argument("--train-api", default="trainer", choices=["trainer", "autograd"],
                    help="C4: nerf.Trainer (default) or the reference-style autograd loop")
    ap.add_argument("--workload", default="render", choices=["render", "train", "c3", "c5"],
                    help="render: BASELINE config 2 (the headline metric); train: BASELINE config 4; "
                         "c3 / c5: the Dex-depth and IR render configurations (secondary lines)")
    args = ap.parse_args()
    if args.workload in SCENES:
        apply_scene(args.workload)
        args.no_train = True
    if args.impl == "reference":
        return run_reference(args)

    import nerf
    from nerf import _lib as L
    if args.precision:
        nerf.set_precision(args.precision)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    dev = torch.device("cuda", local)
    if args.workload == "train":
        run_train(args, dist, rank, world, dev)
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    from nerf.sharding import row_block
    row0, rows = row_block(H, rank, n_gpus)
    mc, mf = state_dicts()
    mc, mf = mc.to(dev), mf.to(dev)
    cfg = make_cfg(nerf)
    ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
    T_cpu, K_cpu = camera()
    T_dev, K_dev = T_cpu.to(dev), K_cpu.to(dev)
    T_pin, K_pin = T_cpu.pin_memory(), K_cpu.pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    from nerf import render as RD

    def step(T_, K_):
        # the public API: nerf.render_camera = get_ray_bundle + run_one_iter_of_nerf for one camera, rays generated
        # inside the setup launch (6 launches per frame, all of them this library's kernels)
        with torch.no_grad():
            return nerf.render_camera(H, W, T_, K_, mc, mf, cfg, mode="validation", encode_position_fn=ex,
                                      encode_direction_fn=ed, m_thres_cand=THRESHOLDS, row_start=row0, row_count=rows)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        step(T_dev, K_dev)
        flush.zero_()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    # per-launch CUDA events recorded by the library on the launching stream inside the timed region
    ev_pool = [RD.Events() for _ in range(args.steps)]
    ev_used = []

    def hook(n):
        e = ev_pool[len(ev_used) % len(ev_pool)]
        ev_used.append((e, n))
        return e
    RD.event_hook = hook
    launches0 = L.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(T_dev, K_dev)
        flush.zero_()
    e1.record()
    barrier()
    RD.event_hook = None
    launches = L.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = reduce_max(e0.elapsed_time(e1) / args.steps)
    per_launch = {}
    for e, n_ in ev_used:
        for name, ms in e.elapsed_ms().items():
            per_launch.setdefault(name, []).append(ms)
    per_launch = {k: sum(v) / len(v) for k, v in per_launch.items()}
    n_rays_rank = rows * W
    kern_ms = per_launch["mlp_fine"]
    kern_name = "mlp_tc" if nerf.get_precision() == "bf16" else "mlp_simt"
    all_mlp_ms = per_launch["mlp_fine"] + per_launch["mlp_coarse"]
    fine = [(kern_ms, n_rays_rank, NC + NF, kern_name)]
    # algorithmic HBM bytes of the small kernels: compositing 24 S + 36 + 4 T per ray in its
    # full form (coarse: weights out, no Dex planes; fine: Dex planes, no weights), resampling 8 Nc + 4 (Nc + Nf)
    hbm_bytes = {"composite_coarse": n_rays_rank * (NC * (16 + 4 + 4) + 12 + 24),
                 "composite_fine": n_rays_rank * ((NC + NF) * (16 + 4) + 12 + 24 + 4 * len(THRESHOLDS)),
                 "resample_merge": n_rays_rank * (8 * NC + 4 * (NC + NF)),
                 "ray_setup": n_rays_rank * (36 + 4 * NC)}
    hbm_log = [(k, per_launch[k], hbm_bytes[k]) for k in hbm_bytes if k in per_launch]

    # ------------------------------------------------------------------ end to end (host buffers)
    out_pin = None
    def e2e_step():
        nonlocal out_pin
        res = step(T_pin.to(dev, non_blocking=True), K_pin.to(dev, non_blocking=True))
        if out_pin is None:
            out_pin = [torch.empty(r.shape, dtype=r.dtype).pin_memory() for r in res]
        for dst, src in zip(out_pin, res):
            dst.copy_(src, non_blocking=True)
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = reduce_max(e0.elapsed_time(e1) / args.steps)
    d2h = sum(o.numel() * o.element_size() for o in out_pin)
    h2d = (T_pin.numel() + K_pin.numel()) * 4

    # ... and with the frame ASSEMBLED in one process (N > 1): every rank's row block is all-gathered over
    # NVLink (nerf.gather_rows: rgb + the T Dex planes + expected depth + acc of the fine pass, what a consumer
    # of the frame reads) and rank 0 copies the full planes to pinned host memory
    ms_gathered, gathered_bytes = None, None
    ms_shared = None
    if dist is not None:
        from nerf.sharding import gather_rows
        full_pin = None
        # (a) nerf.SharedFrame: every rank's compositing kernel writes its rows straight into rank 0's frame over
        # NVLink peer memory - no collective; rank 0 copies the assembled planes to pinned host memory
        try:
            shared = nerf.SharedFrame(H, W, len(THRESHOLDS))
        except L.DexNerfError:
            shared = None
        if shared is not None:
            shared_pin = torch.empty(shared.planes.shape, dtype=torch.float32).pin_memory() if rank == 0 else None
            row0_, rows_ = nerf.row_block(H, rank, world)

            def shared_step():
                nerf.render_camera(H, W, T_pin.to(dev, non_blocking=True), K_pin.to(dev, non_blocking=True), mc, mf, cfg,
                                   mode="validation", encode_position_fn=ex, encode_direction_fn=ed,
                                   m_thres_cand=THRESHOLDS, row_start=row0_, row_count=rows_, frame=shared)
                shared.wait()
                if rank == 0:
                    shared_pin.copy_(shared.planes, non_blocking=True)
            for _ in range(2):
                shared_step()
            barrier()
            e0.record()
            for _ in range(args.steps):
                shared_step()
            e1.record()
            barrier()
            ms_shared = reduce_max(e0.elapsed_time(e1) / args.steps)

        def gathered_step():
            nonlocal full_pin
            res = step(T_pin.to(dev, non_blocking=True), K_pin.to(dev, non_blocking=True))
            planes = torch.cat([res[3]] + [r.unsqueeze(-1) for r in res[4:]], dim=-1)     # (rows, W, 3 + 2 + T)
            full = gather_rows(planes, H)
            if rank == 0:
                if full_pin is None:
                    full_pin = torch.empty(full.shape, dtype=full.dtype).pin_memory()
                full_pin.copy_(full, non_blocking=True)
            return full
        for _ in range(2):
            gathered_step()
        barrier()
        e0.record()
        for _ in range(args.steps):
            full = gathered_step()
        e1.record()
        barrier()
        ms_gathered = reduce_max(e0.elapsed_time(e1) / args.steps)
        gathered_bytes = full.numel() * 4
        assert full.shape[0] == H

    # BASELINE config 4 rides along (a few training iterations; reported under "train_c4")
    train_line = None
    if not args.no_train:
        del flush
        torch.cuda.empty_cache()
        targs = argparse.Namespace(steps=5, warmup=3, train_api=args.train_api)
        train_line = run_train(targs, dist, rank, world, dev, quiet=True)

    if rank == 0:
        pk = peaks()
        rays = H * W
        kern_rays = rows
        flop_launch = float(fine[0][1]) * fine[0][2] * FLOP_PER_EVAL
        achieved = flop_launch / (kern_ms * 1e-3) / 1e12
        peak = pk["tf_sustained"]
        line = {
            "metric": METRIC, "value": rays / (ms_dev * 1e-3),
            "unit": "rays/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if kern_name == "mlp_tc" else "f32", "data": "synthetic",
            "config": workload_config(n_gpus),
            "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": "rays/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": kern_name + " (fine pass, %d samples/ray)" % (NC + NF),
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": pk["source"] + " bf16 sustained",
                         "traffic": ncu_traffic("mlp_tc_fine_c2", fine[0][1]) if SCENE == "c2" else None,
                         "kernel_ms": kern_ms, "mlp_share_of_step": all_mlp_ms / ms_dev,
                         "flop_per_launch": flop_launch},
            "clocks": clocks,
        }
        # the HBM-bound kernels of the step: achieved GB/s of algorithmic bytes against the measured copy peak
        line["hbm_kernels"] = {
            k: {"ms": ms, "bytes": nbytes, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"]} for k, ms, nbytes in hbm_log}
        line["kernel_ms_per_launch"] = {k: round(v, 4) for k, v in sorted(per_launch.items())}
        if ms_shared is not None:
            line["e2e_shared_frame"] = {"value": rays / (ms_shared * 1e-3), "unit": "rays/s", "ms_per_step": ms_shared,
                                        "frame_bytes": H * W * (5 + len(THRESHOLDS)) * 4,
                                        "what": "nerf.SharedFrame: the fine pass's planes written by every rank's "
                                                "compositing kernel straight into rank 0's frame over NVLink peer "
                                                "memory (no collective), barrier, rank 0 copies the frame to pinned "
                                                "host memory"}
        if ms_gathered is not None:
            line["e2e_gathered"] = {"value": rays / (ms_gathered * 1e-3), "unit": "rays/s", "ms_per_step": ms_gathered,
                                    "gathered_bytes_per_step": gathered_bytes,
                                    "what": "fine rgb + expected depth + acc + T Dex planes all-gathered to every rank "
                                            "(NCCL), rank 0 copies the assembled frame to pinned host memory"}
        if train_line is not None:
            line["train_c4"] = {k: train_line[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "gpu_launches",
                                                           "param_checksum", "non_mlp_ms", "small_kernel_ms")}
            line["train_c4"]["kernel_ms"] = train_line["roofline"]["kernel_ms"]
            line["train_c4"]["hbm_frac_of_peak"] = train_line["roofline"].get("frac")
            line["train_c4"]["tensor_frac_of_sustained_peak"] = train_line["roofline"].get("tensor_frac_of_sustained_peak")
        if not args.no_cpu_baseline:
            base = cpu_baseline(reps=3, warmup=1)
            base.pop("ms", None)
            line["cpu_baseline"] = base
            if n_gpus == 1:
                # the informative competitor: the same reference code as eager torch kernels on this GPU
                torch.cuda.empty_cache()
                line["ref_gpu_eager"] = gpu_eager_baseline()
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
