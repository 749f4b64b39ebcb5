"""CPU, world_size 2 (gloo): the image-sharding + single all-gather logic of the multi-GPU path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from depthmap_b200.dist import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 33, 256):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmap_b200.dist import all_gather_batch, shard_range
    full = torch.arange(n_items * 6, dtype=torch.int32).reshape(n_items, 2, 3)
    full16 = (torch.arange(n_items * 4, dtype=torch.int32).reshape(n_items, 4) * 997 % 65536).to(torch.int16).view(torch.uint16)
    lo, hi = shard_range(n_items, rank, world)
    g = all_gather_batch(full[lo:hi].clone(), n_items)
    g16 = all_gather_batch(full16[lo:hi].clone(), n_items)
    ok = torch.equal(g, full) and torch.equal(g16.view(torch.int16), full16.view(torch.int16)) and g16.dtype == torch.uint16
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [8, 5])
def test_all_gather_batch_world2(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _video_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmap_b200.dist import shard_range
    from depthmap_b200.video_mode import exchange_halo, halo_plan
    n_total = 5                                     # rank 0: frames 0-2, rank 1: frames 3-4 (and a 1-frame block case below)
    frames = torch.arange(n_total * 6, dtype=torch.float32).reshape(n_total, 2, 3)
    ok = True
    for total in (5, 3):
        lo, hi = shard_range(total, rank, world)
        before, after = exchange_halo(frames[lo:hi].clone(), None, lo, total)
        bi, ai = halo_plan(lo, hi - lo, total)
        ok = ok and before.shape[0] == len(bi) and after.shape[0] == len(ai)
        ok = ok and all(torch.equal(before[k], frames[g]) for k, g in enumerate(bi)) and all(torch.equal(after[k], frames[g]) for k, g in enumerate(ai))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_video_halo_exchange_world2():
    """sharded video normalisation: every rank ends up with the (at most two) frames before and after its block"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_video_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _round_robin_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmap_b200.dist import all_gather_round_robin, round_robin_owner
    ok = True
    for n in (5, 4, 1, 0):                          # BOOST's patch exchange: uneven deals, a rank without patches, no patch at all
        full32 = torch.arange(n * 7, dtype=torch.float32).reshape(n, 7) * 0.5 + 1.0
        full64 = (torch.arange(n * 3, dtype=torch.float64).reshape(n, 3) + 0.25) ** 2
        mine = list(range(rank, n, world))
        ok = ok and all(round_robin_owner(i, world) == (rank, k) for k, i in enumerate(mine))
        g32 = all_gather_round_robin(full32[mine].clone() if mine else full32[:0].clone(), n)
        g64 = all_gather_round_robin(full64[mine].clone() if mine else full64[:0].clone(), n)
        ok = ok and g32.dtype == torch.float32 and g64.dtype == torch.float64 and torch.equal(g32, full32) and torch.equal(g64, full64)
        # the order-dependent blend every rank replays after the exchange: the same sequence of updates on every rank
        acc = torch.zeros(7, dtype=torch.float32)
        for i in range(n):
            acc = acc * 0.75 + g32[i] * float(g64[i, 0])
        ref = torch.zeros(7, dtype=torch.float32)
        for i in range(n):
            ref = ref * 0.75 + full32[i] * float(full64[i, 0])
        ok = ok and torch.equal(acc, ref)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_boost_patch_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_round_robin_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
