"""Synthetic workload of BASELINE.json: seed-fixed iid reference, reads sampled from it.

Everything is generated on the GPU with torch (plumbing only) straight into the data formats of
the hot path's boundary:

* reference: NGM's own packing of the concatenated reference -- 4 bit/base, high nibble first,
  A0 T1 G2 C3 N4, 1000 N in front of / between / after the contigs (SequenceProvider.cpp:72-109,
  289-330) -- i.e. the body of ``<ref>-enc.2.ngm``;
* reads: ``qry_max_len``-byte NUL-padded upper-case ASCII rows (MappedRead.cpp:25-29);
* candidates: (window_start = loc - corridor/2, read, strand) descriptors, ``cand_begin`` offsets
  per read -- what CS::SendToBuffer hands to ScoreBuffer::addRead (CS.cpp:320-335).

Candidate search itself (CS.cpp) is out of scope, so candidates are synthesised:
the true locus with +-2 bp jitter (bin resolution) plus, for half of the reads, one decoy locus
(CMR/R ~ 1.5).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

SPACER = 1000          # SequenceProvider.cpp:289,321
ASCII_OF_NGM = torch.tensor([ord(c) for c in "ATGCN"] + [ord("N")] * 11, dtype=torch.uint8)


@dataclass
class Reference:
    packed: torch.Tensor        # uint8 [concat_len / 2] on device, NGM packing
    concat_len: int
    contig_start: np.ndarray    # int64 [n_contigs], concatenated coordinates
    contig_len: int


@dataclass
class ReadBatch:
    reads: torch.Tensor         # uint8 [n, qml] ASCII, NUL padded
    pairs: torch.Tensor         # uint8 view of ngm_b200_pair records [n_pairs, 16]
    cand_begin: torch.Tensor    # int32 [n + 1]
    true_pos: torch.Tensor      # int64 [n] concatenated coordinate of the fragment start
    reverse: torch.Tensor       # bool [n]
    n_reads: int
    n_pairs: int


def make_reference(device, n_contigs: int, contig_len: int, seed: int) -> Reference:
    """iid uniform ACGT contigs, packed on the fly (never materialised as ASCII)."""
    assert contig_len % 2 == 0
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    half = contig_len // 2
    total = (SPACER + n_contigs * (contig_len + SPACER)) // 2
    packed = torch.empty(total, dtype=torch.uint8, device=device)
    packed[: SPACER // 2] = 0x44
    starts = []
    off = SPACER // 2
    for _ in range(n_contigs):
        starts.append(off * 2)
        hi = torch.randint(0, 4, (half,), dtype=torch.uint8, device=device, generator=g)
        lo = torch.randint(0, 4, (half,), dtype=torch.uint8, device=device, generator=g)
        packed[off: off + half] = (hi << 4) | lo
        del hi, lo
        off += half
        packed[off: off + SPACER // 2] = 0x44
        off += SPACER // 2
    assert off == total
    return Reference(packed, total * 2, np.array(starts, dtype=np.int64), contig_len)


def _base_codes(ref: Reference, pos: torch.Tensor) -> torch.Tensor:
    """NGM 4-bit codes at concatenated positions `pos` (int64 tensor)."""
    b = ref.packed[pos >> 1]
    return torch.where((pos & 1) == 1, b & 0xF, b >> 4)


def make_reads(ref: Reference, n: int, read_len: int, qml: int, corridor: int, seed: int, sub_rate: float = 0.01,
               indel_rate: float = 0.0005, decoy_fraction: float = 0.5, chunk: int = 1 << 20, paired: bool = False, insert_mean: float = 400.0,
               insert_sd: float = 40.0) -> ReadBatch:
    """paired: rows 2f, 2f + 1 are the mates of fragment f (BASELINE configs[2]; SURVEY 8d: insert ~ N(400, 40), FR orientation, the
    fragment on either strand); true_pos / reverse describe each mate."""
    dev = ref.packed.device
    assert not paired or (n % 2 == 0 and chunk % 2 == 0)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n_contigs = len(ref.contig_start)
    cstart = torch.from_numpy(ref.contig_start).to(dev)
    reads = torch.zeros((n, qml), dtype=torch.uint8, device=dev)
    true_pos = torch.empty(n, dtype=torch.int64, device=dev)
    reverse = torch.empty(n, dtype=torch.bool, device=dev)
    ascii_lut = ASCII_OF_NGM.to(dev)
    p_event = 1.0 - (1.0 - indel_rate) ** read_len
    ar = torch.arange(read_len, device=dev, dtype=torch.int64)[None, :]
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        contig = torch.randint(0, n_contigs, (m,), device=dev, generator=g)
        local = torch.randint(0, ref.contig_len - read_len - 8, (m,), device=dev, generator=g)
        pos = cstart[contig] + local
        pair_rev = None
        if paired:
            fcontig = contig[0::2]
            ins = (torch.randn(m // 2, device=dev, generator=g) * insert_sd + insert_mean).long().clamp(min=read_len + 5, max=4 * int(insert_mean))
            flocal = torch.minimum(local[0::2], ref.contig_len - 8 - ins)
            fpos = cstart[fcontig] + flocal
            flip = torch.rand(m // 2, device=dev, generator=g) < 0.5          # fragment from the minus strand: the mates swap ends
            mate = torch.arange(m, device=dev) & 1
            left_end = mate == flip.repeat_interleave(2).long()
            pos = torch.where(left_end, fpos.repeat_interleave(2), (fpos + ins - read_len).repeat_interleave(2))
            pair_rev = ~left_end
        # at most one 1-3 bp indel per read (0.05 % per base)
        ev = torch.rand(m, device=dev, generator=g) < p_event
        is_ins = torch.rand(m, device=dev, generator=g) < 0.5
        k = torch.randint(1, 4, (m,), device=dev, generator=g)
        at = torch.randint(5, read_len - 8, (m,), device=dev, generator=g)
        k = torch.where(ev, k, torch.zeros_like(k))[:, None]
        at = at[:, None]
        ins = (is_ins & ev)[:, None]
        dele = (~is_ins & ev)[:, None]
        shift = torch.where(dele & (ar >= at), k, torch.zeros_like(ar)) - torch.where(ins & (ar >= at + k), k, torch.zeros_like(ar))
        codes = _base_codes(ref, pos[:, None] + ar + shift)
        inserted = ins & (ar >= at) & (ar < at + k)
        rnd = torch.randint(0, 4, (m, read_len), dtype=torch.uint8, device=dev, generator=g)
        codes = torch.where(inserted, rnd, codes)
        sub = torch.rand((m, read_len), device=dev, generator=g) < sub_rate
        step = torch.randint(1, 4, (m, read_len), dtype=torch.uint8, device=dev, generator=g)
        codes = torch.where(sub & (codes < 4), (codes + step) & 3, codes)
        rev = torch.rand(m, device=dev, generator=g) < 0.5
        if pair_rev is not None:
            rev = pair_rev
        # minus strand: reverse complement; in NGM's code space A0<->T1, G2<->C3 is code ^ 1
        rc = torch.flip(torch.where(codes < 4, codes ^ 1, codes), dims=[1])
        codes = torch.where(rev[:, None], rc, codes)
        reads[s: s + m, :read_len] = ascii_lut[codes.long()]
        true_pos[s: s + m] = pos
        reverse[s: s + m] = rev
        del codes, rc, rnd, sub, step, shift, inserted
    # candidates
    has_decoy = torch.rand(n, device=dev, generator=g) < decoy_fraction
    counts = 1 + has_decoy.to(torch.int32)
    cand_begin = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    cand_begin[1:] = torch.cumsum(counts, 0)
    n_pairs = int(cand_begin[-1].item())
    start = torch.empty(n_pairs, dtype=torch.int64, device=dev)
    ridx = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    flags = torch.empty(n_pairs, dtype=torch.int32, device=dev)
    first = cand_begin[:-1].long()
    jitter = torch.randint(-2, 3, (n,), device=dev, generator=g)
    start[first] = true_pos + jitter - (corridor >> 1)
    ridx[first] = torch.arange(n, dtype=torch.int32, device=dev)
    flags[first] = reverse.to(torch.int32)
    dsel = first[has_decoy] + 1
    nd = int(dsel.numel())
    dcontig = torch.randint(0, n_contigs, (nd,), device=dev, generator=g)
    dlocal = torch.randint(0, ref.contig_len - read_len - 8, (nd,), device=dev, generator=g)
    start[dsel] = cstart[dcontig] + dlocal - (corridor >> 1)
    ridx[dsel] = torch.arange(n, dtype=torch.int32, device=dev)[has_decoy]
    flags[dsel] = torch.randint(0, 2, (nd,), device=dev, generator=g).to(torch.int32)
    pairs = torch.empty((n_pairs, 16), dtype=torch.uint8, device=dev)
    pairs[:, 0:8] = start.view(torch.uint8).view(n_pairs, 8)
    pairs[:, 8:12] = ridx.view(torch.uint8).view(n_pairs, 4)
    pairs[:, 12:16] = flags.view(torch.uint8).view(n_pairs, 4)
    return ReadBatch(reads, pairs, cand_begin, true_pos, reverse, n, n_pairs)


def shapes_for(read_len: int):
    """ReadProvider.cpp:288,304: qry_max_len = (max | 1) + 1, corridor = int(5 + 0.15 * avg)."""
    return (read_len | 1) + 1, int(5 + 0.15 * read_len)
