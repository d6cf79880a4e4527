"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch / row / hypothesis sharding and the
8-byte best-hypothesis max-allreduce that stands for the NCCL one on the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import psulvsb_b200  # noqa: F401
from psulvsb_b200 import sharding


def test_shard_range_partitions():
    for n in [0, 1, 7, 4096, 5000]:
        for world in [1, 2, 3, 8]:
            parts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_triangular_row_range_balances_pairs():
    n = 100_000
    for world in [2, 4, 8]:
        parts = [sharding.triangular_row_range(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        pairs = [sum(n - 1 - i for i in (b, e - 1)) * (e - b) / 2 for b, e in parts]  # arithmetic series
        assert max(pairs) / min(pairs) < 1.01


def test_pack_best_order():
    a = sharding.pack_best(10, 5)
    b = sharding.pack_best(10, 3)
    c = sharding.pack_best(11, 900)
    assert max(a, b) == b and max(a, b, c) == c          # higher count wins, then the lower id
    assert sharding.unpack_best(c) == (11, 900)
    assert sharding.pack_best(2**31 - 1, 0) < 2**63      # stays positive as int64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # hypothesis sharding: every rank scores its slice; the global best is one 8-byte max-allreduce
        rng = np.random.default_rng(123)
        counts = rng.integers(0, 500, 1000)
        counts[[17, 400, 801]] = 777                       # ties across ranks: the first id must win
        b, e = sharding.shard_range(len(counts), rank, world)
        local = counts[b:e]
        k = int(np.argmax(local))                           # first local maximum
        packed = torch.tensor([sharding.pack_best(int(local[k]), b + k)], dtype=torch.int64)
        sharding.allreduce_best(packed)
        best = sharding.unpack_best(int(packed.item()))
        # row sharding: all-gather of the owned rows' popcounts
        n = 1001
        ranges = [sharding.triangular_row_range(n, r, world) for r in range(world)]
        full = torch.arange(n, dtype=torch.int32) * 3 % 17
        mine = torch.zeros(n, dtype=torch.int32)
        rb, re = ranges[rank]
        mine[rb:re] = full[rb:re]
        got = sharding.allgather_row_counts(mine, n, ranges)
        tmax = sharding.max_over_ranks(float(rank + 1))
        q.put((rank, best, bool(torch.equal(got, full)), tmax))
    finally:
        dist.destroy_process_group()


def _edge_worker(rank, world, port, q):
    """What psulvsb_solve_sharded exchanges, replayed on the CPU over gloo: every rank compacts the edges of ITS rows
    (the library's own row partition), the ranks all-gather their counts and then their blocks of the edge list."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from psulvsb_b200 import capi, synth

        n = 700
        pair = synth.make_pair(n, 0.8, 31)
        pi, pj = O.reduced_set(pair["src"], pair["dst"], 0.1)   # the reference's L_reduced_set, row-major
        rb, re = capi.shard_row_range(n, rank, world)
        mine = (pi >= rb) & (pi < re)
        block = torch.from_numpy(np.stack([pi[mine], pj[mine]], axis=1).astype(np.int64))
        cnt = torch.tensor([block.shape[0]], dtype=torch.int64)
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, cnt)
        counts = [int(c.item()) for c in counts]
        width = max(counts)
        padded = torch.zeros((width, 2), dtype=torch.int64)
        padded[: block.shape[0]] = block
        parts = [torch.zeros_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
        full = torch.cat([p[:c] for p, c in zip(parts, counts)]).numpy()
        ok = bool(np.array_equal(full[:, 0], pi) and np.array_equal(full[:, 1], pj))
        pairs_owned = sum(n - 1 - i for i in range(rb, re))
        q.put((rank, ok, (rb, re), pairs_owned, counts))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_row_sharded_edge_lists_concatenate_to_the_reference_order():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_edge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                       # rank-order concatenation == the global row-major list
    assert res[0][2][0] == 0 and res[0][2][1] == res[1][2][0] and res[1][2][1] == 700
    assert abs(res[0][3] - res[1][3]) < 0.01 * (res[0][3] + res[1][3])  # balanced line-vector counts
    assert res[0][4] == res[1][4]


def test_library_row_partition_matches_the_python_one():
    from psulvsb_b200 import capi

    for n in (7, 700, 5000, 100_000):
        for world in (1, 2, 3, 8):
            parts = [capi.shard_row_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            for r in range(world):
                b, e = sharding.triangular_row_range(n, r, world)
                assert abs(parts[r][0] - b) <= 1 and abs(parts[r][1] - e) <= 1  # (round-half differences only)


def test_world2_gloo_best_and_row_counts():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, best, ok, tmax in res:
        assert best == (777, 17)          # identical on every rank, first best id
        assert ok
        assert tmax == 2.0
