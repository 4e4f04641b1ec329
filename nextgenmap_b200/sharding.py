"""Multi-GPU plumbing of the hot path: reads shard embarrassingly, every GPU holds the
whole packed reference, and the only collectives are one broadcast of the reference at start-up and
one sum of the mapping counters at the end.  No collective sits on the data path.

The reference's own multi-device scheme is thread -> device round-robin inside one process
(CS.cpp:441-453) with shared atomics for the counters (NGM.cpp:172-201); here it is one process per
GPU over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_reads: int, rank: int, world: int, paired: bool = False) -> Tuple[int, int]:
    """Contiguous block of reads for `rank`.  Mates of a pair stay in one shard, like
    GetNextReadBatch forcing even batch sizes (NGM.cpp:237-242)."""
    unit = 2 if paired else 1
    units = n_reads // unit
    lo = units * rank // world * unit
    hi = units * (rank + 1) // world * unit
    if rank == world - 1:
        hi = n_reads
    return lo, hi


def broadcast_reference(packed: torch.Tensor, concat_len: int, src: int = 0) -> Tuple[torch.Tensor, int]:
    """Start-up broadcast of the 4-bit packed reference (SequenceProvider packing).  `packed` may be an
    empty tensor on ranks != src; it is (re)allocated to the right size on the same device."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return packed, concat_len
    meta = torch.tensor([concat_len if dist.get_rank() == src else 0], dtype=torch.int64, device=packed.device)
    dist.broadcast(meta, src)
    concat_len = int(meta.item())
    nbytes = (concat_len + 1) // 2
    if dist.get_rank() != src:
        packed = torch.empty(nbytes, dtype=torch.uint8, device=packed.device)
    dist.broadcast(packed, src)
    return packed, concat_len


def reduce_counters(counters: Dict[str, int], device) -> Dict[str, int]:
    """Sum the mapping counters over ranks (NGM.cpp:172-201: reads, mapped, ...; ScoreBuffer::scoreCount,
    AlignmentBuffer::alignmentCount)."""
    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
