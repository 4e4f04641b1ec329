"""N > 1 host logic on CPU: world_size-2 gloo run of the read sharding, the start-up reference broadcast and
the final counter reduction (the only collectives of the path)."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nextgenmap_b200 import sharding


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(7)
        full = torch.from_numpy(rng.integers(0, 256, 5001, dtype=np.uint8))
        packed = full.clone() if rank == 0 else torch.empty(0, dtype=torch.uint8)
        packed, concat_len = sharding.broadcast_reference(packed, 10001 if rank == 0 else 0, src=0)
        assert concat_len == 10001 and torch.equal(packed, full)
        # the prefix table travels the same way (rank 0 read it from NGM's cache file / exported it from its device)
        tab = rng.integers(0, 1 << 31, 1025, dtype=np.uint32)
        weight = rng.integers(-5, 100, 1025).astype(np.int8)
        table = rng.integers(0, 1 << 32, 7777, dtype=np.uint64).astype(np.uint32)
        got = sharding.broadcast_index((tab, weight, table) if rank == 0 else None, "cpu", src=0)
        assert all(np.array_equal(g, w) and g.dtype == w.dtype for g, w in zip(got, (tab, weight, table)))
        n_reads = 1001
        lo, hi = sharding.shard_range(n_reads, rank, world)
        # every rank "maps" its shard: pretend reads with an index divisible by 7 stay unmapped
        idx = np.arange(lo, hi)
        counters = sharding.reduce_counters({"reads": hi - lo, "mapped": int(np.count_nonzero(idx % 7)), "pairs_scored": int(idx.sum() % 1000)}, "cpu")
        t = sharding.max_over_ranks(10.0 + rank, "cpu")
        out.put((rank, lo, hi, counters, t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_roundtrip():
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, c0, t0), (r1, lo1, hi1, c1, t1) = res
    assert (lo0, hi1) == (0, 1001) and hi0 == lo1                       # shards tile the read set
    assert c0 == c1                                                     # all ranks see the reduced counters
    assert c0["reads"] == 1001 and c0["mapped"] == int(np.count_nonzero(np.arange(1001) % 7))
    assert t0 == t1 == 11.0                                             # timing is the max over ranks


def test_shard_range_keeps_mates_together_and_covers_everything():
    for n, world in [(10, 3), (1001, 8), (7, 8), (20_000_000, 8)]:
        spans = [sharding.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    spans = [sharding.shard_range(1000, r, 3, paired=True) for r in range(3)]
    assert all(lo % 2 == 0 for lo, _ in spans) and spans[-1][1] == 1000
