"""world_size-2 checks of the multi-GPU host logic on CPU with the gloo backend: the row-block
partition the render bench uses, the gather of finished planes, and the flat gradient all-reduce
of the data-parallel training step."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf.sharding import allreduce_gradients, gather_rows, row_block


def test_row_blocks_tile_the_image():
    for H in (800, 720, 270, 7, 1):
        for world in (1, 2, 3, 4, 8):
            blocks = [row_block(H, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == H
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # render sharding: every rank fills its rows with a function of the global pixel index
        start, count = row_block(H, rank, world)
        rows = torch.arange(start, start + count, dtype=torch.float32)[:, None, None]
        local = rows * 1000 + torch.arange(W, dtype=torch.float32)[None, :, None] + torch.tensor([0.0, 0.25, 0.5])
        full = gather_rows(local, H)
        ref = (torch.arange(H, dtype=torch.float32)[:, None, None] * 1000
               + torch.arange(W, dtype=torch.float32)[None, :, None] + torch.tensor([0.0, 0.25, 0.5]))
        assert torch.equal(full, ref)
        # training: flat all-reduce == mean of the per-rank gradients
        torch.manual_seed(0)
        a, b = torch.nn.Linear(5, 3), torch.nn.Linear(3, 2)
        for i, p in enumerate(list(a.parameters()) + list(b.parameters())):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        n = allreduce_gradients([a, b])
        assert n == sum(p.numel() for p in list(a.parameters()) + list(b.parameters()))
        mean = sum(range(1, world + 1)) / world
        for i, p in enumerate(list(a.parameters()) + list(b.parameters())):
            assert torch.allclose(p.grad, torch.full_like(p, mean * (i + 1)))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 9, 4), nprocs=2, join=True)
