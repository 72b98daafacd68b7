"""Multi-GPU: the path shards by position (one position = one independent unit;
NN_output[i] depends only on input i, tree_builder.pyx:318).  One process per GPU
(torch.distributed for the plumbing), weights replicated, leaf batches split into
contiguous slices, NO collective on the data path.  The only communication is an optional
end-of-run gather of a few int64 counters (NCCL on GPUs, gloo in CPU tests)."""


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n units owned by `rank` of `world` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tree_owner(tree_id, world):
    """Config 5 (independent game trees): tree t is served by GPU t mod world."""
    return int(tree_id) % int(world)


def gather_counters(counters, device=None):
    """all_gather a short list of int64 counters (positions, ns, top-1 hits ...) across ranks.
    Returns a [world, len(counters)] int64 tensor on every rank; works without an initialised
    process group (world = 1)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(c) for c in counters], dtype=torch.int64, device=device or "cpu")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None].clone()
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)
