"""Multi-GPU plumbing for the scoring path: variants / mutants / sequences are independent units, so
ranks take contiguous index ranges (output order is preserved by concatenation) and there is no
collective on the data path.  A caller that wants the whole score matrix on one rank (or on all)
gathers the per-rank tiles once at the end with a single NCCL collective over NVLink
(SURVEY §8e).  Works with any ``torch.distributed`` backend (NCCL on GPUs, gloo in the CPU tests).
"""


def shard_range(n, rank, world):
    """Contiguous range [lo, hi) of rank ``rank`` out of ``world`` over ``n`` units: the first
    ``n % world`` ranks get one extra unit."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n, world):
    return [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]


def gather_tiles(local, n_total, dst=None, group=None, out=None):
    """Gathers per-rank score tiles ``local`` (n_local, F) into the (n_total, F) matrix with ONE collective.

    ``dst=None``: every rank receives the full matrix (all_gather); otherwise only rank ``dst`` does
    and the others get ``None`` (gather).  Ranks receive straight into row blocks of one output tensor
    (``out``, optional, (world * max shard, F), reusable across calls): no staging copies.  Shards may differ
    by one row; they are padded to the largest shard for the collective and the result is compacted."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    assert local.shape[0] == sizes[rank], "local tile does not match this rank's shard"
    m = max(sizes)
    F = local.shape[1]
    padded = local
    if local.shape[0] != m:
        padded = torch.zeros((m, F), dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
    padded = padded.contiguous()
    receives = dst is None or rank == dst
    if receives and (out is None or tuple(out.shape) != (world * m, F) or out.dtype != local.dtype):
        out = torch.empty((world * m, F), dtype=local.dtype, device=local.device)
    if dst is None:
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        blocks = [out[r * m:(r + 1) * m] for r in range(world)] if rank == dst else None
        dist.gather(padded, blocks, dst=dst, group=group)
        if rank != dst:
            return None
    if all(s == m for s in sizes):
        return out
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)], dim=0)
