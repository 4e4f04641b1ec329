"""Host-side mirror of NextGenMap's mapping run on top of the device library:

    reads -> CS (k-mer vote, CS.cpp) -> ScoreBuffer (BatchScore of every candidate; top1SE + computeMQ, or top1PE for pairs)
          -> AlignmentBuffer (BatchAlign of the winner) -> GenericReadWriter filters -> SAMWriter lines

Everything numeric runs in ``libngm_b200.so`` (there is no CPU path); this module only strings the device calls together
and renders the records the way ``SAMWriter::DoWriteReadGeneric`` / ``DoWriteUnmappedReadGeneric`` / ``DoWritePair`` do
(src/writer/SAMWriter.cpp:98-228,230-310,312-365), after the post-processing of ``AlignmentBuffer::DoRun`` / ``WriteRead``
(src/AlignmentBuffer.cpp:121-135,166-208) and the output filters of ``GenericReadWriter::WriteRead`` / ``WritePair``
(src/writer/GenericReadWriter.h:190-312).  ``topn`` 1, no bs-mapping -- what BASELINE configs[0]-[2] run.

The renderers take any object with the fields of ``MappedBatch`` (the tests also feed them from the CPU oracles to pin those against
the unmodified NextGenMap).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

import ctypes as C

from .cuda_sw import ALIGN_REC, PAIR, CudaSW, MapResult, SamBatch, SamOpts, load_library

COMP = bytes.maketrans(b"ACGT", b"TGCA")
U64 = 2 ** 64 - 1


@dataclass
class MappedBatch:
    """Per-read results of one batch (arrays of length n_reads unless noted)."""
    cand_begin: np.ndarray      # int32 [n + 1]
    pairs: np.ndarray           # PAIR [total]  candidates in the reference's order
    scores: np.ndarray          # float32 [total]  BatchScore
    max_hit: np.ndarray         # float32  MappedRead::s
    best_pair: np.ndarray       # int32  index into pairs, -1 = no candidate handed to alignment
    mapq: np.ndarray            # int32
    num_top: np.ndarray         # int32  MappedRead::numTopScores
    recs: np.ndarray            # ALIGN_REC fields position_offset, qstart, qend, nm, identity, score
    heap: Optional[np.ndarray]  # uint8 string heap (CIGAR, MD)
    strings: Callable[[int], Tuple[bytes, bytes]] = None      # read -> (CIGAR, MD bytes as written)
    pair_fail: Optional[np.ndarray] = None                    # int32  NGMNames::PairedFail set by top1PE (paired runs)


def _search_and_score(sw: CudaSW, reads: np.ndarray, mode: int):
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    begin, pairs, votes, max_hit = sw.cs_search(reads)
    sw.set_reads(reads)
    scores = sw.score_pairs(mode, pairs)
    return reads, begin, pairs, scores, max_hit


def _align_winners(sw: CudaSW, mode: int, n: int, pairs: np.ndarray, best: np.ndarray):
    winners = np.zeros(n, dtype=PAIR)
    has = best >= 0
    winners[has] = pairs[best[has]]
    winners["read_index"] = np.arange(n, dtype=np.uint32)
    winners["flags"][~has] = 4                     # NGM_B200_PAIR_SKIP
    recs, heap = sw.align_pairs(mode, winners)
    raw = heap.tobytes()

    def strings(r: int):
        o, cl, ml = int(recs[r]["str_off"]), int(recs[r]["cigar_len"]), int(recs[r]["md_len"])
        return raw[o: o + cl], raw[o + cl: o + cl + ml]
    return recs, heap, strings


def map_reads(sw: CudaSW, reads: np.ndarray, mode: int = 0) -> MappedBatch:
    """One single-end batch through candidate search, scoring, top-1 selection and alignment.  Needs set_reference + a prefix table."""
    import torch
    reads, begin, pairs, scores, max_hit = _search_and_score(sw, reads, mode)
    n = reads.shape[0]
    dev = torch.device("cuda", sw.params.device)
    st = torch.cuda.current_stream(dev).cuda_stream
    d_begin = torch.from_numpy(begin).to(dev)
    d_scores = torch.from_numpy(scores).to(dev) if len(scores) else torch.zeros(1, dtype=torch.float32, device=dev)
    d_best = torch.empty(n, dtype=torch.int32, device=dev)
    d_mq = torch.empty(n, dtype=torch.int32, device=dev)
    d_nt = torch.empty(n, dtype=torch.int32, device=dev)
    sw._check(sw.lib.ngm_b200_dev_select_top1_ex(sw.ctx, n, d_begin.data_ptr(), d_scores.data_ptr(), d_best.data_ptr(), d_mq.data_ptr(), d_nt.data_ptr(), st))
    torch.cuda.synchronize(dev)
    best = d_best.cpu().numpy()
    recs, heap, strings = _align_winners(sw, mode, n, pairs, best)
    return MappedBatch(begin, pairs, scores, max_hit, best, d_mq.cpu().numpy(), d_nt.cpu().numpy(), recs, heap, strings)


def map_pairs(sw: CudaSW, reads: np.ndarray, mode: int = 0) -> MappedBatch:
    """One paired-end batch (rows 2f, 2f + 1 = the mates of fragment f; ``sw.pe_configure`` once per run before the first batch):
    candidate search and scoring per mate, top1PE / top1SE on the device, alignment of every selected candidate."""
    import torch
    reads, begin, pairs, scores, max_hit = _search_and_score(sw, reads, mode)
    n = reads.shape[0]
    dev = torch.device("cuda", sw.params.device)
    st = torch.cuda.current_stream(dev).cuda_stream
    d_begin = torch.from_numpy(begin).to(dev)
    d_scores = torch.from_numpy(scores).to(dev) if len(scores) else torch.zeros(1, dtype=torch.float32, device=dev)
    d_pairs = torch.from_numpy(pairs.view(np.uint8).reshape(-1, 16)).to(dev) if len(pairs) else torch.zeros((1, 16), dtype=torch.uint8, device=dev)
    out = [torch.empty(n, dtype=torch.int32, device=dev) for _ in range(4)]
    sw._check(sw.lib.ngm_b200_dev_select_pairs(sw.ctx, n, d_begin.data_ptr(), d_pairs.data_ptr(), d_scores.data_ptr(), len(pairs), out[0].data_ptr(),
                                               out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), st))
    torch.cuda.synchronize(dev)
    best, mq, nt, pf = (t.cpu().numpy() for t in out)
    recs, heap, strings = _align_winners(sw, mode, n, pairs, best)
    return MappedBatch(begin, pairs, scores, max_hit, best, mq, nt, recs, heap, strings, pf)


def map_reads_topn(sw: CudaSW, reads: np.ndarray, topn: int, strata: bool = False, mode: int = 0):
    """Single-end batch with ``topn`` > 1 (NGM ``-n``): candidate search, scores, ScoreBuffer::topNSE on the device, up to topn alignments
    per read.  -> the batch ``sam_lines_topn`` takes."""
    import torch
    from types import SimpleNamespace
    reads, begin, pairs, scores, max_hit = _search_and_score(sw, reads, mode)
    n = reads.shape[0]
    dev = torch.device("cuda", sw.params.device)
    st = torch.cuda.current_stream(dev).cuda_stream
    d_begin = torch.from_numpy(begin).to(dev)
    d_scores = torch.from_numpy(scores).to(dev) if len(scores) else torch.zeros(1, dtype=torch.float32, device=dev)
    d_sel = torch.empty((n, topn), dtype=torch.int32, device=dev)
    d_ns, d_mq, d_nt = (torch.empty(n, dtype=torch.int32, device=dev) for _ in range(3))
    sw._check(sw.lib.ngm_b200_dev_select_topn(sw.ctx, n, d_begin.data_ptr(), d_scores.data_ptr(), len(scores), topn, 1 if strata else 0, d_sel.data_ptr(),
                                              d_ns.data_ptr(), d_mq.data_ptr(), d_nt.data_ptr(), st))
    torch.cuda.synchronize(dev)
    sel = d_sel.cpu().numpy()
    rr, jj = np.nonzero(sel >= 0)
    winners = pairs[sel[rr, jj]].copy()
    recs = np.zeros((n, topn), dtype=ALIGN_REC)
    recs["score"] = -1.0
    raw, heap = b"", np.zeros(1, np.uint8)
    if len(rr):
        flat, heap = sw.align_pairs(mode, winners)
        recs[rr, jj] = flat
        raw = heap.tobytes()

    def strings(r: int, j: int):
        o, cl, ml = int(recs[r, j]["str_off"]), int(recs[r, j]["cigar_len"]), int(recs[r, j]["md_len"])
        return raw[o: o + cl], raw[o + cl: o + cl + ml]
    return SimpleNamespace(cand_begin=begin, pairs=pairs, scores=scores, max_hit=max_hit, sel=sel, n_sel=d_ns.cpu().numpy(), mapq=d_mq.cpu().numpy(),
                           num_top=d_nt.cpu().numpy(), recs=recs, strings=strings, heap=heap)


def map_batch(sw: CudaSW, reads: np.ndarray, mode: int = 0, paired: bool = False, capacity: int = 0, heap_bytes: int = 0):
    """The same as ``map_reads`` / ``map_pairs`` / ``map_reads_topn`` through the one-call entry point ``ngm_b200_map_batch`` (host buffers
    in, host buffers out; candidates, scores and winners never leave the device in between).  After ``sw.se_configure(topn > 1)`` a
    single-end batch comes back in the shape ``map_reads_topn`` returns (recs [n, topn], sel, n_sel)."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n, stride = reads.shape
    topn = 1 if paired else getattr(sw, "se_topn", 1)
    cap, scap = capacity or max(1024, 4 * n), heap_bytes or n * 64 * topn + 4096
    for _ in range(4):
        begin = np.zeros(n + 1, np.int32)
        pairs, scores = np.zeros(cap, dtype=PAIR), np.zeros(cap, np.float32)
        best, mq, nt, pf = (np.zeros(n, np.int32) for _ in range(4))
        mh = np.zeros(n, np.float32)
        recs, heap = np.zeros((n, topn) if topn > 1 else n, dtype=ALIGN_REC), np.zeros(scap, np.uint8)
        sel, n_sel = (np.zeros((n, topn), np.int32), np.zeros(n, np.int32)) if topn > 1 else (None, None)
        res = MapResult(begin.ctypes.data, pairs.ctypes.data, scores.ctypes.data, cap, 0, best.ctypes.data, mq.ctypes.data, nt.ctypes.data,
                        pf.ctypes.data if paired else None, mh.ctypes.data, recs.ctypes.data, heap.ctypes.data, scap, 0,
                        sel.ctypes.data if topn > 1 else None, n_sel.ctypes.data if topn > 1 else None)
        rc = sw.lib.ngm_b200_map_batch(sw.ctx, reads.ctypes.data, n, stride, mode, 1 if paired else 0, C.byref(res))
        if rc == -3:                                 # (the library puts the insert-size sums of a paired run back before it reports this)
            cap = max(cap, int(res.n_candidates) + 16)
            scap = max(scap, int(res.str_used) + 16)
            continue
        sw._check(rc)
        total, raw = int(res.n_candidates), heap.tobytes()      # (the heap is sparse: one slot per sub-batch of the library's pipeline)
        if topn > 1:
            from types import SimpleNamespace

            def strings_n(r: int, j: int):
                o, cl, ml = int(recs[r, j]["str_off"]), int(recs[r, j]["cigar_len"]), int(recs[r, j]["md_len"])
                return raw[o: o + cl], raw[o + cl: o + cl + ml]
            return SimpleNamespace(cand_begin=begin, pairs=pairs[:total], scores=scores[:total], max_hit=mh, sel=sel, n_sel=n_sel, mapq=mq, num_top=nt,
                                   recs=recs, strings=strings_n, heap=heap, best_pair=best)

        def strings(r: int):
            o, cl, ml = int(recs[r]["str_off"]), int(recs[r]["cigar_len"]), int(recs[r]["md_len"])
            return raw[o: o + cl], raw[o + cl: o + cl + ml]
        return MappedBatch(begin, pairs[:total], scores[:total], mh, best, mq, nt, recs, heap, strings, pf if paired else None)
    raise RuntimeError("ngm_b200_map_batch: buffer sizing failed")


def _xi(identity: float) -> str:
    # SAMWriter.cpp:188-189: round(Identity * 10000.0f) / 10000.0f, printed with %g
    v = float(np.float32(identity) * np.float32(10000.0))
    r = math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)
    return "%g" % float(np.float32(r / 10000.0))


@dataclass
class _Read:
    """One read after AlignmentBuffer::DoRun: what WriteRead / WritePair look at."""
    name: str
    seq: bytes
    qual: bytes
    length: int
    has: bool = False           # hasCandidates() with a usable alignment
    reverse: bool = False
    loc: int = 0                # Location.m_Location (concatenated until convert() succeeds)
    contig: int = 0             # Location.getrefId()
    converted: bool = False
    bp: int = -1
    r: int = -1


def _collect(batch, reads: np.ndarray, names, quals, encref, min_mq: int = 0) -> List[_Read]:
    out = []
    for r in range(reads.shape[0]):
        seq = reads[r].tobytes().split(b"\0")[0]
        rd = _Read(names[r], seq, quals[r], len(seq), r=r)
        bp = int(batch.best_pair[r])
        # AlignmentBuffer::addRead (AlignmentBuffer.cpp:46-49): a read below min_mq is written as unmapped without being aligned
        if bp >= 0 and float(batch.recs[r]["score"]) >= 0.0 and int(batch.mapq[r]) >= min_mq:
            p = batch.pairs[bp]
            rd.has, rd.bp = True, bp
            rd.reverse = bool(int(p["flags"]) & 1)
            # AlignmentBuffer.cpp:129: Location += PositionOffset - (corridor >> 1); window_start is Location - corridor/2 already
            rd.loc = (int(p["window_start"]) + int(batch.recs[r]["position_offset"])) & U64
            conv = encref.convert(rd.loc)               # AlignmentBuffer.cpp:171-175
            if conv is not None:
                rd.contig, rd.loc, rd.converted = conv[0], conv[1], True
        out.append(rd)
    return out


def _passes(batch, rd: _Read, min_identity: float, min_residues: float) -> bool:
    rec = batch.recs[rd.r]
    mres = rd.length * min_residues if min_residues <= 1.0 else min_residues      # GenericReadWriter.h:206-210,273-278
    return bool(float(rec["identity"]) >= np.float32(min_identity)
                and float(rd.length - int(rec["qstart"]) - int(rec["qend"])) >= np.float32(mres))


def _mapped_line(batch, rd: _Read, encref, flags: int, rnext: str, pnext: int, tlen: int) -> str:
    """SAMWriter::DoWriteReadGeneric (SAMWriter.cpp:98-228)."""
    rec = batch.recs[rd.r]
    cigar, md = batch.strings(rd.r)
    s, q = rd.seq, rd.qual
    if rd.reverse:
        s = rd.seq.translate(COMP)[::-1]
        q = rd.qual if rd.qual[:1] == b"*" else rd.qual[::-1]      # SAMWriter.cpp:122: a quality string starting with "*" is left alone
        flags |= 0x10
    ntop = int(batch.num_top[rd.r])
    qstart, qend = int(rec["qstart"]), int(rec["qend"])
    zs = []
    if getattr(batch, "bs_mapping", 0) == 1:                    # SAMWriter.cpp:173-187 (set ``batch.bs_mapping = 1`` for a --bs-mapping run)
        second = getattr(batch, "pair_fail", None) is not None and (rd.r & 1)
        zs = ["ZS:Z:" + (("+-" if rd.reverse else "--") if second else ("-+" if rd.reverse else "++"))]
    return "\t".join([rd.name, str(flags), encref.contigs[rd.contig][0], str((rd.loc + 1) & 0xFFFFFFFF), str(int(batch.mapq[rd.r])), cigar.decode(),
                      rnext, str((pnext + 1) & 0xFFFFFFFF), str(tlen), s.decode(), q.decode(), "AS:i:%d" % int(batch.scores[rd.bp]),
                      "NM:i:%d" % int(rec["nm"]), "NH:i:%d" % ntop, *zs, "XI:f:" + _xi(float(rec["identity"])), "X0:i:%d" % ntop,
                      "XE:i:%d" % int(batch.max_hit[rd.r]), "XR:i:%d" % (rd.length - qstart - qend), "MD:Z:" + md.split(b"\0")[0].decode(),
                      *(_slam_tags(cigar, md.split(b"\0")[0], s, qstart, rd.reverse) if getattr(batch, "slam_seq", 0) else [])])


_TRANS = {c: i for i, cs in enumerate((b"Aa", b"Cc", b"Gg", b"Tt")) for c in cs}


def _slam_tags(cigar: bytes, md: bytes, oriented: bytes, qstart: int, reverse: bool) -> List[str]:
    """TC:i / RA:Z / MP:Z of a --slam-seq run (set ``batch.slam_seq``): GenericReadWriter::computeSlaSeqTags (GenericReadWriter.h:87-181) over the
    AlignmentPosition list computeCigarMD keeps (SWOclCigar.cpp:484-497,523-535), rebuilt from CIGAR + MD + the read as it was aligned."""
    import re
    rates = [0] * 25
    mp = []
    read_pos, ref_pos = qstart, 0
    toks = re.findall(rb"\d+|\^[A-Za-z]+|[A-Za-z]", md)        # numbers, deletions, single mismatch letters
    ti, eq = 0, 0
    for ln, op in re.findall(rb"(\d+)([MIDSH=X])", cigar):
        ln = int(ln)
        if op in b"SH":
            continue
        if op == b"I":
            read_pos += ln
        elif op == b"D":
            if ti < len(toks) and toks[ti].isdigit():
                ti += 1
            ti += 1                                              # the ^... token
            ref_pos += ln
        else:
            for _ in range(ln):
                q = _TRANS.get(oriented[read_pos], 4)
                while eq == 0 and ti < len(toks) and toks[ti].isdigit():
                    eq = int(toks[ti])
                    ti += 1
                if eq > 0:
                    eq -= 1
                    rates[5 * q + q] += 1
                else:
                    t = 5 * _TRANS.get(toks[ti][0], 4) + q
                    ti += 1
                    rates[t] += 1
                    mp.append("%d:%d:%d" % (t, read_pos + 1, ref_pos + 1))
                read_pos += 1
                ref_pos += 1
    out = ["TC:i:%d" % (rates[2] if reverse else rates[16]), "RA:Z:" + ",".join(str(v) for v in rates)]
    if mp:
        out.append("MP:Z:" + ",".join(mp))
    return out


def _i32(v: int) -> int:
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _unmapped_line(rd: _Read, flags: int, rname: str = "*", loc: int = -1, rnext: str = "*", pnext: int = -1) -> str:
    """SAMWriter::DoWriteUnmappedReadGeneric (SAMWriter.cpp:312-365)."""
    return "\t".join([rd.name, str(flags | 0x4), rname, str(_i32(loc + 1)), "0", "*", rnext, str(_i32(pnext + 1)), "0", rd.seq.decode(), rd.qual.decode()])


def sam_lines(sw: Optional[CudaSW], batch, reads: np.ndarray, names: Sequence[str], quals: Sequence[bytes], encref, corridor: int,
              min_identity: float = 0.65, min_residues: float = 0.5, min_mq: int = 0) -> List[str]:
    """SAM body lines of a single-end batch (no header), one per read, in read order."""
    out = []
    for rd in _collect(batch, reads, names, quals, encref, min_mq):
        if rd.has and rd.converted and _passes(batch, rd, min_identity, min_residues):
            out.append(_mapped_line(batch, rd, encref, 0, "*", -1, 0))
        else:
            out.append(_unmapped_line(rd, 0))
    return out


def sam_lines_topn(batch, reads: np.ndarray, names: Sequence[str], quals: Sequence[bytes], encref, corridor: int, min_identity: float = 0.65,
                   min_residues: float = 0.5) -> List[str]:
    """Single-end run with ``topn`` > 1: up to topn lines per read.  ``batch``: sel [n, topn] (candidate indices in the order of
    ScoreBuffer::topNSE, -1 padded), n_sel, mapq, num_top, recs [n, topn], strings(r, j).  AlignmentBuffer::WriteRead converts every
    location and keeps the LAST result (AlignmentBuffer.cpp:169-175); GenericReadWriter::WriteRead (GenericReadWriter.h:200-248) stops
    accepting alignments after the first one that fails a filter, drops repeated locations, and writes the unmapped record when nothing
    passed; every line after the best candidate's carries 0x100 (SAMWriter.cpp:116-118)."""
    out = []
    for r in range(reads.shape[0]):
        seq = reads[r].tobytes().split(b"\0")[0]
        length, ns = len(seq), int(batch.n_sel[r])
        lines, mapped, seen, views = [], ns > 0, set(), []
        for j in range(ns):
            p = batch.pairs[int(batch.sel[r, j])]
            rec = batch.recs[r, j]
            rd = _Read(names[r], seq, quals[r], length, has=True, reverse=bool(int(p["flags"]) & 1), bp=int(batch.sel[r, j]), r=r)
            rd.loc = (int(p["window_start"]) + int(rec["position_offset"])) & U64
            conv = encref.convert(rd.loc)
            if conv is not None:
                rd.contig, rd.loc, rd.converted = conv[0], conv[1], True
            mapped = conv is not None                      # the last convert() decides
            views.append((rd, rec))
        if any(float(rec["score"]) < 0.0 for _, rec in views):
            mapped = False
        for j, (rd, rec) in enumerate(views):
            mres = length * min_residues if min_residues <= 1.0 else min_residues
            mapped = mapped and bool(float(rec["identity"]) >= np.float32(min_identity)) and bool(float(length - int(rec["qstart"]) - int(rec["qend"])) >= np.float32(mres))
            if mapped:
                key = (rd.loc, rd.contig, rd.reverse)
                if key not in seen:
                    seen.add(key)
                    one = SimpleBatch(batch, r, j)
                    lines.append(_mapped_line(one, rd, encref, 0x100 if j else 0, "*", -1, 0))
        out += lines if lines else [_unmapped_line(_Read(names[r], seq, quals[r], length), 0)]
    return out


class SimpleBatch:
    """View of alignment j of read r in a topn batch with the field names ``_mapped_line`` reads."""

    def __init__(self, batch, r: int, j: int):
        self.recs = {r: batch.recs[r, j]}
        self.mapq, self.num_top, self.max_hit, self.scores = batch.mapq, batch.num_top, batch.max_hit, batch.scores
        self.strings = lambda rr: batch.strings(rr, j)


def sam_lines_paired(batch, reads: np.ndarray, names: Sequence[str], quals: Sequence[bytes], encref, corridor: int, min_identity: float = 0.65,
                     min_residues: float = 0.5, min_mq: int = 0, min_insert_size: int = 0, max_insert_size: int = 1000) -> List[str]:
    """SAM body lines of a paired batch, two per fragment: AlignmentBuffer::WriteRead's pair check (AlignmentBuffer.cpp:176-200),
    GenericReadWriter::WritePair's filters (GenericReadWriter.h:258-312) and SAMWriter::DoWritePair (SAMWriter.cpp:230-310)."""
    rds = _collect(batch, reads, names, quals, encref, min_mq)
    if max_insert_size <= 0:
        max_insert_size = 2 ** 31 - 1
    out = []
    for f in range(0, len(rds) - 1, 2):
        a, b = rds[f], rds[f + 1]                       # a: first mate (ReadId even), b: second mate
        fail = bool(batch.pair_fail[a.r]) or bool(batch.pair_fail[b.r])
        # WriteRead: the mate that arrives second does the check; with both aligned that is the first mate (top1PE submits read = second
        # mate first), so read = a, read->Paired = b
        if a.has and b.has:
            d = (b.loc - a.loc + a.length) if b.loc > a.loc else (a.loc - b.loc + b.length)
            d = _i32(d)
            if a.contig != b.contig or d < min_insert_size or d > max_insert_size or a.reverse == b.reverse:
                fail = True
        for rd in (a, b):                               # WritePair: mapped1 / mapped2, clearScores() otherwise
            if rd.has and not _passes(batch, rd, min_identity, min_residues):
                rd.has = False
        fa, fb = 0x1 | 0x40, 0x1 | 0x80
        ra, rb = encref.contigs[a.contig][0], encref.contigs[b.contig][0]
        if not a.has and not b.has:
            out += [_unmapped_line(b, fb | 0x8), _unmapped_line(a, fa | 0x8)]
        elif not a.has:
            out += [_mapped_line(batch, b, encref, fb | 0x8, "=", b.loc, 0), _unmapped_line(a, fa, rb, b.loc, "=", b.loc)]
        elif not b.has:
            out += [_unmapped_line(b, fb, ra, a.loc, "=", a.loc), _mapped_line(batch, a, encref, fa | 0x8, "=", a.loc, 0)]
        elif not fail:
            fa |= 0x2
            fb |= 0x2
            fwd, rev = (a, b) if not a.reverse else (b, a)          # after the check above exactly one mate is on the minus strand
            rec = batch.recs[rev.r]
            dist = _i32((rev.loc + rev.length - int(rec["qstart"]) - int(rec["qend"])) - fwd.loc)
            ffwd, frev = (fa, fb) if fwd is a else (fb, fa)
            out += [_mapped_line(batch, rev, encref, frev, "=", fwd.loc, -dist), _mapped_line(batch, fwd, encref, ffwd | 0x20, "=", rev.loc, dist)]
        else:
            if a.reverse:
                fb |= 0x20
            if b.reverse:
                fa |= 0x20
            out += [_mapped_line(batch, b, encref, fb, ra, a.loc, 0), _mapped_line(batch, a, encref, fa, rb, b.loc, 0)]
    return out


def format_sam(batch, reads: np.ndarray, names: Sequence[str], quals: Sequence[bytes], encref, paired: bool, min_identity: float = 0.65,
               min_residues: float = 0.5, min_insert_size: int = 0, max_insert_size: int = 1000, threads: int = 0, min_mq: int = 0,
               clip_seq: bool = False, read_group: Optional[str] = None, bs_mapping: int = 0, slam_seq: int = 0) -> bytes:
    """The same lines as ``sam_lines`` / ``sam_lines_paired`` from the library's multi-threaded formatter (``ngm_b200_format_sam``): what a
    C / C++ host calls.  ``encref``: an ``EncodedReference`` (the C struct is handed over as it is).  ``batch.recs`` / ``batch.heap`` as
    ``ngm_b200_align_pairs`` returned them."""
    lib = load_library()
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n, stride = reads.shape
    q = np.zeros((n, stride), np.uint8)
    for i, ql in enumerate(quals):
        q[i, : len(ql)] = np.frombuffer(ql, np.uint8)
    name_arr = (C.c_char_p * n)(*[nm.encode() for nm in names])
    topn = int(batch.sel.shape[1]) if hasattr(batch, "sel") else 0        # a ``map_reads_topn`` batch: recs [n, topn], sel, n_sel
    best = batch.sel[:, 0] if topn else batch.best_pair
    keep = [np.ascontiguousarray(batch.pairs), np.ascontiguousarray(batch.scores, dtype=np.float32), np.ascontiguousarray(best, dtype=np.int32),
            np.ascontiguousarray(batch.mapq, dtype=np.int32), np.ascontiguousarray(batch.num_top, dtype=np.int32),
            np.ascontiguousarray(batch.pair_fail, dtype=np.int32) if paired else None, np.ascontiguousarray(batch.max_hit, dtype=np.float32),
            np.ascontiguousarray(batch.recs), np.ascontiguousarray(batch.heap)]
    assert keep[7].dtype == ALIGN_REC
    ptr = lambda a: None if a is None else a.ctypes.data
    sel = np.ascontiguousarray(batch.sel, dtype=np.int32) if topn else None
    n_sel = np.ascontiguousarray(batch.n_sel, dtype=np.int32) if topn else None
    sb = SamBatch(n, stride, reads.ctypes.data, q.ctypes.data, name_arr, ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), ptr(keep[4]), ptr(keep[5]),
                  ptr(keep[6]), ptr(keep[7]), ptr(keep[8]), topn if topn > 1 else 0, ptr(sel), ptr(n_sel))
    so = SamOpts(min_identity, min_residues, min_insert_size, max_insert_size, threads, min_mq, 1 if clip_seq else 0, read_group.encode() if read_group else None,
                 bs_mapping, slam_seq)
    used = C.c_size_t(0)
    cap = n * max(topn, 1) * (2 * stride + 256) + 4096
    for _ in range(2):
        out = np.zeros(cap, np.uint8)
        rc = lib.ngm_b200_format_sam(C.byref(encref.c), C.byref(so), C.byref(sb), out.ctypes.data, cap, C.byref(used))
        if rc == -3:
            cap = used.value + 16
            continue
        if rc < 0:
            raise RuntimeError(f"ngm_b200_format_sam failed ({rc})")
        return out[: used.value].tobytes()
    raise RuntimeError("ngm_b200_format_sam: buffer sizing failed")
