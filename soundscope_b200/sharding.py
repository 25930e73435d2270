"""Multi-GPU plumbing for the analyzer path: streams are independent (every meter's state is private to
its `EbuR128`, reference src/analyzer.rs:29-32), so a batch is partitioned into contiguous blocks of
streams, one block per rank, with no data-path exchange.  The only collective is the gather of the
per-stream result rows (ssb_results_device: [momentary, shortterm, integrated, LRA, true_peak[C],
sample_peak[C]]).  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests.
"""


def shard_range(n_streams, world, rank):
    """Contiguous block [lo, hi) of rank's streams; blocks differ in size by at most one stream."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    lo = (n_streams * rank) // world
    hi = (n_streams * (rank + 1)) // world
    return lo, hi


def shard_sizes(n_streams, world):
    return [shard_range(n_streams, world, r)[1] - shard_range(n_streams, world, r)[0] for r in range(world)]


def gather_results(local_rows, n_streams, group=None):
    """all_gather of [n_local, stride] result rows into [n_streams, stride] in global stream order.

    Ranks may own different numbers of streams; rows are padded to the largest shard for the collective
    and trimmed afterwards.  Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(n_streams, world)
    assert local_rows.shape[0] == sizes[dist.get_rank(group)], "local rows do not match this rank's shard"
    stride = local_rows.shape[1]
    pad = max(sizes)
    buf = local_rows
    if local_rows.shape[0] != pad:
        buf = torch.zeros((pad, stride), dtype=local_rows.dtype, device=local_rows.device)
        buf[: local_rows.shape[0]] = local_rows
    out = torch.empty((world * pad, stride), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if all(s == pad for s in sizes):
        return out
    return torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)], dim=0)


class PeerGather:
    """The gather of a BatchAnalyzer's result rows over all ranks, for one rank.

    kind "p2p": every results launch of the analyzer also stores this rank's rows into every rank's gather buffer
    (CUDA IPC peer memory over NVLink, include/soundscope_b200.h "multi-GPU") — no collective kernel; `wait()`
    enqueues the arrival check and flips the double buffer.  kind "nccl": the fallback when peer mapping is not
    available — one all_gather_into_tensor of the latest rows per `wait()`.

        g = PeerGather(an, world, rank)            # collective: every rank constructs it
        an.add_frames_results_device(x, g.local_rows())    # any number of times
        rows = g.wait()                            # [world * n_streams, stride], rows of every rank's latest launch
    """

    def __init__(self, an, env=None, world=None, rank=None, group=None, allow_p2p=True):
        import torch
        import torch.distributed as dist
        self.an, self.group = an, group
        self.world = world if world is not None else (env.world if env is not None else dist.get_world_size(group))
        self.rank = rank if rank is not None else (env.rank if env is not None else dist.get_rank(group))
        self._local = torch.empty((an.n_streams, an.stride), dtype=torch.float64, device="cuda")
        self.parity = 0
        self.kind = "nccl all_gather_into_tensor per wait()"
        self._all = None
        ok = 0
        if allow_p2p and self.world <= 8:
            try:
                handle = an.gather_create(self.world, self.rank)
                ok = 1
            except Exception:
                handle = bytes(64)
        else:
            handle = bytes(64)
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, (ok, handle), group=group)
        else:
            handles = [(ok, handle)]
        if all(h[0] for h in handles):
            try:
                an.gather_open(b"".join(h[1] for h in handles))
                opened = 1
            except Exception:
                opened = 0
        else:
            opened = 0
        if self.world > 1:
            flags = [None] * self.world
            dist.all_gather_object(flags, opened, group=group)
        else:
            flags = [opened]
        if all(flags):
            self.kind = "p2p: result rows stored into every rank's buffer by the results epilogue (CUDA IPC over NVLink), no collective kernel"
            self._p2p = True
            an.gather_select(0)
        else:
            self._p2p = False
            if ok:
                an.gather_destroy()
            self._all = torch.empty((self.world * an.n_streams, an.stride), dtype=torch.float64, device="cuda")

    def local_rows(self):
        return self._local

    def publish(self):
        """p2p: the results launch has already published; nccl: the rows travel in wait()."""
        return None

    def wait(self):
        """Rows of every rank's latest results launch, [world * n_streams, stride], valid in stream order."""
        import torch.distributed as dist
        if self._p2p:
            self.an.gather_wait()
            rows = self.an.gather_rows(self.parity, self.world)
            self.parity ^= 1
            self.an.gather_select(self.parity)   # the next launches fill the other half while this one is read
            return rows
        dist.all_gather_into_tensor(self._all, self._local, group=self.group)
        return self._all

    def close(self):
        if self._p2p:
            self.an.sync()
            self.an.gather_destroy()
            self._p2p = False
