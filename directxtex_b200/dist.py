"""Multi-GPU plumbing for batches: one process per GPU, contiguous image ranges per rank, no
data-path collective during compute, ONE all-gather of the packed BC blocks at the end.  torch.distributed is
plumbing only (NCCL on GPUs; gloo in the CPU tests)."""


def shard_range(n_items, world, rank):
    """contiguous range [lo, hi) of items owned by `rank`: rank r owns [r*n/G, (r+1)*n/G)"""
    lo = (rank * n_items) // world
    hi = ((rank + 1) * n_items) // world
    return lo, hi


def max_shard(n_items, world):
    return max(shard_range(n_items, world, r)[1] - shard_range(n_items, world, r)[0] for r in range(world))


def all_gather_blocks(local_blocks, n_items, item_bytes, world, rank, dist, torch):
    """local_blocks: uint8 tensor with this rank's packed blocks (its shard's items back to back).
    Returns a uint8 tensor with all n_items*item_bytes bytes in item order on every rank.
    Uneven shards are padded to the largest shard so that a single all_gather_into_tensor suffices."""
    per = max_shard(n_items, world) * item_bytes
    send = local_blocks
    if send.numel() != per:
        send = torch.zeros(per, dtype=torch.uint8, device=local_blocks.device)
        send[: local_blocks.numel()] = local_blocks
    recv = torch.empty(per * world, dtype=torch.uint8, device=local_blocks.device)
    dist.all_gather_into_tensor(recv, send)
    if n_items % world == 0:
        return recv
    out = torch.empty(n_items * item_bytes, dtype=torch.uint8, device=local_blocks.device)
    for r in range(world):
        lo, hi = shard_range(n_items, world, r)
        out[lo * item_bytes: hi * item_bytes] = recv[r * per: r * per + (hi - lo) * item_bytes]
    return out
