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


class _DeviceMemory:
    """A raw device allocation as a __cuda_array_interface__ object (float32 vector)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class SharedFrame:
    """The frame of a row-sharded render ASSEMBLED in one process without a collective: rank `root` holds the planes of
    the fine pass - rgb (H, W, 3), expected depth (H, W), accumulation (H, W) and the T Dex-NeRF depth planes (T, H, W) -
    in an IPC-exported allocation (dexnerf_p2p_alloc) that every rank maps (cudaIpcOpenMemHandle), and every rank's
    final compositing kernel writes its row block STRAIGHT into them: the stores travel over NVLink while the kernel
    runs, there is no all-gather, no padding to equal blocks and no concatenation afterwards.

        frame = nerf.SharedFrame(H, W, len(thresholds))            # once, collectively
        row0, rows = nerf.row_block(H, rank, world)
        nerf.render_camera(H, W, pose, K, mc, mf, cfg, ..., row_start=row0, row_count=rows, frame=frame)
        frame.wait()                                               # every rank's block has landed
        if rank == frame.root: use(frame.rgb, frame.depth, frame.acc, frame.dex)

    One process per GPU on one node with peer access; otherwise construction raises DexNerfError (use gather_rows)."""

    def __init__(self, height, width, n_thresholds, group=None, root=0, device=None):
        import ctypes as C
        import os
        import torch.distributed as dist
        from . import _lib as L
        self.H, self.W, self.T = int(height), int(width), int(n_thresholds)
        self.group, self.root = group, int(root)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        hw = self.H * self.W
        self.numel = hw * (3 + 1 + 1 + self.T)
        lib = L.lib()
        me = self.device.index
        devs = [None] * self.world
        dist.all_gather_object(devs, (os.uname().nodename, me), group=group)
        ok = len({d[0] for d in devs}) == 1 and len({d[1] for d in devs}) == self.world and \
            all(d[1] == devs[self.root][1] or torch.cuda.can_device_access_peer(d[1], devs[self.root][1]) for d in devs)
        if not ok:
            raise L.DexNerfError("SharedFrame needs one process per GPU on one node with peer access to the root's GPU")
        self._own = self._mapped = None
        handle = [None]
        if self.rank == self.root:
            ptr = C.c_void_p()
            L.check(lib.dexnerf_p2p_alloc(self.numel * 4, C.byref(ptr)), "p2p_alloc")
            h = C.create_string_buffer(64)
            L.check(lib.dexnerf_p2p_export(ptr, h), "p2p_export")
            self._own, base, handle = ptr.value, ptr.value, [h.raw]
        dist.broadcast_object_list(handle, src=dist.get_global_rank(group, self.root) if group is not None else self.root,
                                   group=group)
        opened = True
        if self.rank != self.root:
            q = C.c_void_p()
            if lib.dexnerf_p2p_open(C.create_string_buffer(handle[0], 64), C.byref(q)) == 0:
                self._mapped, base = q.value, q.value
            else:
                opened = False
        everyone = [None] * self.world            # all ranks succeed or all give up (nobody may wait alone later)
        dist.all_gather_object(everyone, opened, group=group)
        if not all(everyone):
            if self._mapped:
                lib.dexnerf_p2p_close(self._mapped)
            dist.barrier(group=group)
            if self._own:
                lib.dexnerf_p2p_free(self._own)
            self._own = self._mapped = None
            raise L.DexNerfError("SharedFrame: a rank could not map the root's frame (CUDA IPC): "
                                 + lib.dexnerf_last_error().decode())
        flat = torch.as_tensor(_DeviceMemory(base, self.numel), device=self.device)
        self.rgb = flat[:3 * hw].view(self.H, self.W, 3)
        self.depth = flat[3 * hw:4 * hw].view(self.H, self.W)
        self.acc = flat[4 * hw:5 * hw].view(self.H, self.W)
        self.dex = flat[5 * hw:].view(self.T, self.H, self.W) if self.T else None
        self.planes = flat                     # rgb | depth | acc | dex, the root's copy source

    def outputs(self, row_start, row_count):
        """[rgb, depth, acc, dex] destinations of the rows [row_start, row_start + row_count) for the fused render
        call (dex: the first plane's block; the planes are H * W floats apart)."""
        r0, r1 = int(row_start), int(row_start) + int(row_count)
        dex = self.dex[0, r0:r1].reshape(-1) if self.T else None
        return [self.rgb[r0:r1].reshape(-1, 3), self.depth[r0:r1].reshape(-1), self.acc[r0:r1].reshape(-1), dex,
                self.H * self.W]

    def wait(self):
        """Returns (stream-ordered) once every rank's kernels that wrote into the frame have finished."""
        import torch.distributed as dist
        dist.barrier(group=self.group)

    def close(self):
        import torch.distributed as dist
        from . import _lib as L
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.rgb = self.depth = self.acc = self.dex = self.planes = None
        if self._mapped:
            L.lib().dexnerf_p2p_close(self._mapped)
        dist.barrier(group=self.group)
        if self._own:
            L.lib().dexnerf_p2p_free(self._own)
        self._own = self._mapped = None
