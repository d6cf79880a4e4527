"""Multi-GPU sharding of the PSULVSB hot path (one process per GPU, torch.distributed).

The path shards three natural ways, none of which moves point data:
  * independent fragment pairs (batched registration): a contiguous slice of the batch per rank,
    no data-path collective at all;
  * consistency-mask row blocks (large N): each rank builds the mask rows it owns
    (psulvsb_consistency_mask_rows); only the per-row popcounts (4 B per row) are all-gathered so
    every rank knows the global reduced-set size;
  * hypothesis batches (scoring sweep): each rank scores its slice with psulvsb_score_batch and the
    global best is ONE 8-byte max-allreduce of (count << 32) | (0xFFFFFFFF - hypothesis id).
The helpers below are backend-agnostic: NCCL on the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n items for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank / world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def triangular_row_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Row block [begin, end) of the upper-triangular pair set such that every rank owns ~the same
    number of pairs (row i has n - 1 - i of them): boundaries at n (1 - sqrt(1 - k / world))."""
    def bound(k: int) -> int:
        if k <= 0:
            return 0
        if k >= world:
            return n
        return int(round(n * (1.0 - (1.0 - k / world) ** 0.5)))
    return bound(rank), bound(rank + 1)


def pack_best(count: int, hyp_id: int) -> int:
    """(count << 32) | (0xFFFFFFFF - id): max() picks the highest count, then the LOWEST id
    (the reference keeps the first best hypothesis: strict '>' at registration.cc:1337)."""
    return (int(count) << 32) | (0xFFFFFFFF - int(hyp_id))


def unpack_best(packed: int) -> tuple[int, int]:
    packed &= (1 << 64) - 1
    return packed >> 32, 0xFFFFFFFF - (packed & 0xFFFFFFFF)


def allreduce_best(packed: torch.Tensor) -> torch.Tensor:
    """In-place max-allreduce of the packed 64-bit best-hypothesis key (int64 tensor of 1 element).
    Counts are < 2^31, so the signed int64 order equals the unsigned one."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.MAX)
    return packed


def allgather_row_counts(local_counts: torch.Tensor, n: int, ranges) -> torch.Tensor:
    """Every rank contributes the popcounts of the rows it owns; returns the full [n] vector.
    `ranges`: list of (begin, end) per rank.  local_counts has n entries (only the owned ones valid)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_counts
    world = dist.get_world_size()
    rank = dist.get_rank()
    width = max(e - b for b, e in ranges)
    mine = torch.zeros(width, dtype=local_counts.dtype, device=local_counts.device)
    b, e = ranges[rank]
    mine[: e - b] = local_counts[b:e]
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = torch.zeros(n, dtype=local_counts.dtype, device=local_counts.device)
    for (bb, ee), p in zip(ranges, parts):
        out[bb:ee] = p[: ee - bb]
    return out


def max_over_ranks(value: float, device=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
