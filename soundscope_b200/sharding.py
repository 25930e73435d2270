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
