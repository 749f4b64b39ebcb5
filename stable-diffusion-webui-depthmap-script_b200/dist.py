"""Multi-GPU plumbing: images are independent units, so a batch shards by image across ranks (one process per GPU)
and the finished tensors are exchanged with ONE all-gather per output tensor (NCCL over NVLink on the GPU box; gloo in
the CPU tests).  No collective exists inside the operators themselves."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of n_items owned by `rank`; sizes differ by at most one, earlier ranks get the extras."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_batch(local, n_items: int, group=None):
    """Gather per-rank shards (dim 0, contiguous shard_range slices of an n_items batch) into the full batch on every
    rank.  Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    assert local.shape[0] == hi - lo, "local shard does not match shard_range"
    pad = local
    if hi - lo < max_n:
        pad = torch.cat([local, local.new_zeros((max_n - (hi - lo),) + tuple(local.shape[1:]))], dim=0)
    pad = pad.contiguous()
    # move raw bytes: every backend has uint8 collectives, not every backend has uint16 / int16 ones
    row_bytes = pad[0].numel() * pad.element_size() if max_n > 0 else 0
    view = pad.view(torch.uint8).reshape(max_n, row_bytes)
    out = view.new_empty((world * max_n, row_bytes))
    dist.all_gather_into_tensor(out, view, group=group)
    parts = [out[r * max_n: r * max_n + (sizes[r][1] - sizes[r][0])] for r in range(world)]
    full = torch.cat(parts, dim=0).contiguous()
    return full.view(local.dtype).reshape((n_items,) + tuple(local.shape[1:]))


def round_robin_owner(i: int, world: int):
    """BOOST deals patch i to rank i % world, which keeps it in slot i // world (patches are sorted largest first, so a round-robin
    deal balances the work better than contiguous slices)."""
    return i % world, i // world


def all_gather_round_robin(local, n_items: int, group=None):
    """local: [k, ...] = this rank's items i = rank, rank + world, ... (in that order) of an n_items list dealt round-robin ->
    [n_items, ...] in item order on every rank, with ONE all-gather (ranks with fewer items are zero-padded to ceil(n / world) slots)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = -(-n_items // world)
    mine = len(range(rank, n_items, world))
    assert local.shape[0] == mine, "local items do not match the round-robin deal"
    send = local.new_zeros((per,) + tuple(local.shape[1:]))
    if mine:
        send[:mine].copy_(local)
    out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    if per:
        dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    idx = [r * per + slot for r, slot in (round_robin_owner(i, world) for i in range(n_items))]
    return out[torch.tensor(idx, dtype=torch.long, device=out.device)] if n_items else out[:0]
