"""Channel sharding across the GPUs of one box.

Channels are independent streams (one reference object per channel, example.cpp:30-41), so the
multi-GPU decomposition is a static partition of the channel axis: rank r owns a contiguous range,
holds its own per-channel state, and NO collective runs in steady state.  torch.distributed (NCCL
on GPUs, gloo in the CPU tests) is used only to move a single [channels, frames] buffer that
lives on rank 0 to/from the shards, and for the barrier / max-reduce of timings in bench.py.
"""


def shard_channels(n_channels, world_size, rank):
    """Contiguous, balanced partition: returns (first_channel, count) of `rank`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(int(n_channels), int(world_size))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def scatter_channels(x_full, n_channels, frames, dist, device=None, dtype=None, src=0):
    """Rank `src` holds x_full [n_channels, frames]; every rank gets its shard [count, frames].
    Point-to-point sends of contiguous row slabs (ncclSend/ncclRecv under NCCL)."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    start, count = shard_channels(n_channels, world, rank)
    if rank == src:
        dtype = x_full.dtype
        device = x_full.device
    mine = torch.empty((count, frames), dtype=dtype, device=device)
    if rank == src:
        reqs = []
        for r in range(world):
            s, c = shard_channels(n_channels, world, r)
            if r == src:
                mine.copy_(x_full[s:s + c])
            elif c > 0:
                reqs.append(dist.isend(x_full[s:s + c].contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif count > 0:
        dist.recv(mine, src=src)
    return mine


def gather_channels(y_shard, n_channels, dist, dst=0):
    """Inverse of scatter_channels: rank `dst` returns [n_channels, frames], others None."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    frames = y_shard.shape[1]
    if rank == dst:
        full = torch.empty((n_channels, frames), dtype=y_shard.dtype, device=y_shard.device)
        for r in range(world):
            s, c = shard_channels(n_channels, world, r)
            if r == dst:
                full[s:s + c].copy_(y_shard)
            elif c > 0:
                buf = torch.empty((c, frames), dtype=y_shard.dtype, device=y_shard.device)
                dist.recv(buf, src=r)
                full[s:s + c].copy_(buf)
        return full
    if y_shard.shape[0] > 0:
        dist.send(y_shard.contiguous(), dst=dst)
    return None
