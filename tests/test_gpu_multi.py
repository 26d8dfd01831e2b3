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
