"""Row-block partition of an image over the ranks of one node.

Rendering is embarrassingly parallel over rays, so each rank renders a contiguous block of image
rows with replicated weights and NO collective on the data path; the finished planes can be
gathered to rank 0 (torch.distributed all_gather over NCCL/NVLink, or gloo on CPU in the tests).
Training is data parallel: every rank draws its own rays and the two MLPs' gradients are summed
with one all-reduce of a flat fp32 buffer, scaled by 1/world so that the result is the gradient
of the mean loss the reference computes over its batch (train_dexnerf_rgb.py:264-281)."""
import torch


def row_block(height, rank, world):
    """(row_start, row_count) of `rank`: blocks differ by at most one row and tile [0, height)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d not in [0, %d)" % (rank, world))
    base, extra = divmod(int(height), int(world))
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def gather_rows(local, height, group=None):
    """All-gather row blocks `local` (rows_r, ...) of every rank into the full (height, ...) tensor.
    Blocks may differ by one row, so they are padded to the largest block for the collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [row_block(height, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: counts[rank]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def allreduce_gradients(models, group=None):
    """Sum the gradients of `models` (the coarse/fine pair) across ranks through ONE flat fp32 buffer
    and scale by 1/world (mean-loss semantics).  Returns the number of floats reduced."""
    import torch.distributed as dist
    grads = [p.grad for m in models if m is not None for p in m.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).to(torch.float32) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / dist.get_world_size(group))
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off
