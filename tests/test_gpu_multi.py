"""Multi-GPU checks that need two real GPUs (skipped otherwise): nerf.Trainer's data-parallel step - the fused
all-reduce + Adam kernel over NVLink peer memory (csrc/p2p.cu, the default) and the overlapped NCCL all-reduce
(DEXNERF_P2P=0) - gives every rank the parameters a single process gets from the concatenated batch, and the two
paths agree with each other."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import torch.distributed as dist
import nerf
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
def make():
    torch.manual_seed(7)
    return nerf.FlexibleNeRFModel(8, 128, 3, 6, 4).cuda(), nerf.FlexibleNeRFModel(8, 128, 3, 6, 4).cuda()
mode = dict(chunksize=1 << 20, perturb=True, num_coarse=16, num_fine=24, white_background=False,
            radiance_field_noise_std=0.2, lindisp=False)
cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=2.0, far=6.0), nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
ex, ed = nerf.get_embedding_function(6, True, True), nerf.get_embedding_function(4, True, True)
g = torch.Generator().manual_seed(5)
n = 64                                   # per rank
N = n * world
ro = (torch.randn(N, 3, generator=g) * 0.2).cuda(); rd = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=-1).cuda()
tgt = torch.rand(N, 3, generator=g).cuda()
rng = dict(t_rand=torch.rand(N, 16, generator=g).cuda(), u=torch.rand(N, 24, generator=g).cuda(),
           noise_coarse=(0.2 * torch.randn(N, 16, generator=g)).cuda(), noise_fine=(0.2 * torch.randn(N, 40, generator=g)).cuda())
sl = slice(rank * n, (rank + 1) * n)
mc, mf = make()
tr = nerf.Trainer(mc, mf, cfg, ex, ed, world_size=world)
want_p2p = os.environ.get("DEXNERF_P2P", "1") != "0"
assert (tr._p2p is not None) == want_p2p, "peer-memory path: expected %s" % want_p2p
for _ in range(3):
    tr.step(ro[sl], rd[sl], tgt[sl], rng={k: v[sl] for k, v in rng.items()})
# single-process reference: the whole batch at once (mean over N rays == mean of the per-rank means)
mc1, mf1 = make()
t1 = nerf.Trainer(mc1, mf1, cfg, ex, ed, world_size=1)
for _ in range(3):
    t1.step(ro, rd, tgt, rng=rng)
diff = (tr.params - t1.params).abs()
frac = float((diff > 2e-4).float().mean())
gathered = [torch.empty_like(tr.params) for _ in range(world)]
dist.all_gather(gathered, tr.params)
same = all(torch.equal(gathered[0], x) for x in gathered)
if rank == 0:
    print("RESULT frac_diff=%.5f ranks_identical=%s sum=%.9e" % (frac, same, float(tr.params.double().sum())))
tr.close()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_trainer_data_parallel_matches_single_process(tmp_path, p2p):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "2954" + p2p, str(script)],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, DEXNERF_P2P=p2p))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    frac = float(line.split("frac_diff=")[1].split()[0])
    assert "ranks_identical=True" in line
    assert frac < 0.02, line          # Adam steps of entries with |grad| ~ eps aside, the update is the same


FRAME_WORKER = r'''
import os, sys, torch
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dex-nerf_b200"))
import torch.distributed as dist
import nerf
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
torch.manual_seed(3)
mc, mf = nerf.FlexibleNeRFModel(8, 128, 3, 10, 4), nerf.FlexibleNeRFModel(8, 128, 3, 10, 4)
with torch.no_grad():
    for m in (mc, mf):
        m.fc_alpha.weight.mul_(300.0)
mc, mf = mc.cuda(), mf.cuda()
mode = dict(chunksize=1 << 20, perturb=False, num_coarse=64, num_fine=64, white_background=False,
            radiance_field_noise_std=0.0, lindisp=False)
cfg = nerf.CfgNode(dict(dataset=dict(no_ndc=True, near=0.3, far=4.0), nerf=dict(use_viewdirs=True, train=dict(mode), validation=dict(mode))))
ex, ed = nerf.get_embedding_function(10, True, True), nerf.get_embedding_function(4, True, True)
H, W = 27, 48                      # 27 rows over 2 ranks: blocks of 14 and 13
K = torch.tensor([[40.0, 0, 24.0], [0, 40.0, 13.5], [0, 0, 1]]).cuda()
T = torch.eye(4); T[2, 3] = 1.5; T = T.cuda()
thr = [float(m) for m in range(5, 105, 5)]
frame = nerf.SharedFrame(H, W, len(thr))
row0, rows = nerf.row_block(H, rank, world)
kw = dict(mode="validation", encode_position_fn=ex, encode_direction_fn=ed, m_thres_cand=thr)
ok = True
with torch.no_grad():
    for it in range(3):            # the frame is reused
        out = nerf.render_camera(H, W, T, K, mc, mf, cfg, row_start=row0, row_count=rows, frame=frame, **kw)
        frame.wait()
        if rank == frame.root:
            full = nerf.render_camera(H, W, T, K, mc, mf, cfg, **kw)          # the whole frame in this process
            ok = ok and torch.equal(frame.rgb, full[3]) and torch.equal(frame.depth, full[4]) and torch.equal(frame.acc, full[5])
            ok = ok and all(torch.equal(frame.dex[t], full[6 + t]) for t in range(len(thr)))
            ok = ok and torch.equal(out[3], full[3][row0:row0 + rows]) and torch.equal(out[6], full[6][row0:row0 + rows])
            crossed = float((torch.stack(full[6:]) > 0.3 + 1e-4).float().mean())
        dist.barrier()             # nobody overwrites the frame while the root compares
if rank == frame.root:
    print("RESULT frame_identical=%s crossed=%.3f" % (ok, crossed))
frame.close()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_shared_frame_assembles_the_row_blocks_over_peer_memory(tmp_path):
    """nerf.SharedFrame: two ranks render their row blocks (14 and 13 rows) with the compositing kernel writing straight
    into rank 0's frame over NVLink peer memory; the assembled planes - rgb, expected depth, accumulation and the 20 Dex
    depth planes - are bit-identical to the same frame rendered in one process, three frames in a row."""
    script = tmp_path / "frame_worker.py"
    script.write_text("ROOT = %r\n" % ROOT + FRAME_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(script)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    assert "frame_identical=True" in line, line
    assert float(line.split("crossed=")[1]) > 0.02, line
