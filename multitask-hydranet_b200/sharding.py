"""Batch sharding for multi-GPU inference: images are independent, so ranks take contiguous slices of
the global batch and never exchange data.  The only collective is the MAX
reduction of the per-rank device time used for reporting."""


def shard_bounds(global_batch, world_size, rank):
    """Contiguous [begin, end) slice of rank ``rank``; sizes differ by at most one image."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(global_batch, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value, dist=None, device=None):
    """MAX all-reduce of a python float (identity when not distributed)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
