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


def broadcast_index(arrays, device, src: int = 0):
    """Start-up broadcast of the prefix table: rank `src` holds the three arrays of NGM's
    `<ref>-ht-<k>-<skip>.3.ngm` (tab uint32 [4^k + 1], weight int8 [4^k + 1], table uint32 [n]) -- read from the cache file or
    exported from a device-built index -- and every other rank receives them (NCCL over NVLink on GPUs) and installs them with
    ``CudaSW.cs_load_index``.  `arrays` is that triple of numpy arrays on `src`, anything (None) elsewhere.  -> the triple on every rank."""
    import numpy as np
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return arrays
    me = dist.get_rank()
    dtypes = (np.uint32, np.int8, np.uint32)
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if me == src:
        meta = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=device)
    dist.broadcast(meta, src)
    out = []
    for i, (n, dt) in enumerate(zip(meta.tolist(), dtypes)):
        nbytes = int(n) * np.dtype(dt).itemsize
        if me == src:
            t = torch.from_numpy(np.ascontiguousarray(arrays[i], dtype=dt).view(np.uint8).reshape(-1)).to(device)
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        dist.broadcast(t, src)
        out.append(arrays[i] if me == src else t.cpu().numpy().view(dt))
        del t
    return tuple(out)


def broadcast_index_device(sw, params, device, src: int = 0) -> dict:
    """The same on the device (multi-GPU runs): rank `src` has an index in `sw` (built on its GPU or loaded from NGM's cache file); it is
    exported into device buffers, broadcast over NCCL / NVLink and installed on every other rank's GPU (ngm_b200_dev_cs_export_index /
    ngm_b200_dev_cs_load_index).  `params`: the CsParams of the run.  -> cs_index_info() of this rank."""
    import ctypes as C
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return sw.cs_index_info()
    me = dist.get_rank()
    meta = torch.zeros(3, dtype=torch.int64, device=device)
    if me == src:
        info = sw.cs_index_info()
        meta = torch.tensor([info["index_len"], info["table_len"], info["max_kfreq"]], dtype=torch.int64, device=device)
    dist.broadcast(meta, src)
    index_len, table_len, max_kfreq = (int(v) for v in meta.tolist())
    d_tab = torch.empty(index_len, dtype=torch.int32, device=device)
    d_weight = torch.empty(index_len, dtype=torch.int8, device=device)
    d_table = torch.empty(max(table_len, 1), dtype=torch.int32, device=device)
    st = torch.cuda.current_stream(device).cuda_stream
    lib = sw.lib
    lib.ngm_b200_dev_cs_export_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_cs_load_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    if me == src:
        sw._check(lib.ngm_b200_dev_cs_export_index(sw.ctx, d_tab.data_ptr(), d_weight.data_ptr(), d_table.data_ptr(), st))
    for t in (d_tab, d_weight, d_table):
        dist.broadcast(t, src)
    if me != src:
        p = type(params)(params.kmer, params.kmer_skip, params.bin_size, params.skip_rep, params.sensitivity, params.kmer_min, max_kfreq, params.max_cmrs)
        sw._check(lib.ngm_b200_dev_cs_load_index(sw.ctx, C.byref(p), d_tab.data_ptr(), d_weight.data_ptr(), index_len, d_table.data_ptr(), table_len, st))
    return sw.cs_index_info()


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
