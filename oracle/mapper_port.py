"""oracle/mapper_port.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

A whole mapping run on the CPU restatements, strung together the way NextGenMap does it:
candidate search (oracle/cs_oracle.c) -> window decode + BatchScore (oracle/ngm_oracle.c) -> selection (oracle/select_oracle.c:
top1SE or top1PE) -> window decode + BatchAlign.  The result has the fields of nextgenmap_b200.host.pipeline.MappedBatch, so the same
SAM renderer serves the device run and this one; tests/test_mapper_oracle.py compares its SAM with the unmodified NextGenMap's,
which pins the selection restatement (there is no finer-grained probe of ScoreBuffer in the reference).
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Optional

import numpy as np

from oracle import cs_port, port

PAIR = np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])
REC = np.dtype([("position_offset", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("nm", "<i4"), ("identity", "<f4"), ("score", "<f4")])
SEL_CAND = np.dtype([("location", "<u8"), ("score", "<f4"), ("orig", "<i4")])
SEL_RESULT = np.dtype([("best", "<i4"), ("mapq", "<i4"), ("num_top", "<i4"), ("paired_fail", "<i4"), ("insert", "<i4")])
COMP = bytes.maketrans(b"ACGT", b"TGCA")
U64 = 2 ** 64 - 1


class _SelParams(C.Structure):
    _fields_ = [("pair_score_cutoff", C.c_float), ("min_insert", C.c_int), ("max_insert", C.c_int), ("strata", C.c_int), ("fast_pairing", C.c_int)]


class _SelState(C.Structure):
    _fields_ = [("dist_sum", C.c_longlong), ("dist_count", C.c_longlong)]


class Selector:
    """ScoreBuffer's selection state for one run with one CS thread (pairDistSum / pairDistCount carried over the batches)."""

    def __init__(self, pair_score_cutoff: float = 0.9, min_insert_size: int = 0, max_insert_size: int = 1000, strata: int = 0, fast_pairing: int = 0):
        self.lib = port.lib()
        self.p = _SelParams(pair_score_cutoff, min_insert_size, max_insert_size if max_insert_size > 0 else 2 ** 31 - 1, strata, fast_pairing)
        self.state = _SelState(0, 1)                          # ScoreBuffer.h:90

    def select_pairs(self, cand_begin: np.ndarray, locations: np.ndarray, scores: np.ndarray, read_len: np.ndarray) -> np.ndarray:
        """-> SEL_RESULT [n_reads]; best indexes the candidate arrays."""
        n = len(cand_begin) - 1
        cands = np.zeros(max(len(scores), 1), dtype=SEL_CAND)
        cands["location"][: len(scores)] = locations
        cands["score"][: len(scores)] = scores
        out = np.zeros(n, dtype=SEL_RESULT)
        begin = np.ascontiguousarray(cand_begin, dtype=np.int32)
        rl = np.ascontiguousarray(read_len, dtype=np.int32)
        self.lib.sel_oracle_select_pairs(n, begin.ctypes.data_as(C.c_void_p), cands.ctypes.data_as(C.c_void_p), rl.ctypes.data_as(C.c_void_p),
                                         C.byref(self.p), C.byref(self.state), out.ctypes.data_as(C.c_void_p))
        return out

    def select_single(self, cand_begin: np.ndarray, scores: np.ndarray) -> np.ndarray:
        n = len(cand_begin) - 1
        out = np.zeros(n, dtype=SEL_RESULT)
        out["best"] = -1
        out["num_top"] = 1
        for r in range(n):
            b, e = int(cand_begin[r]), int(cand_begin[r + 1])
            if e > b:
                cands = np.zeros(e - b, dtype=SEL_CAND)
                cands["score"] = scores[b:e]
                cands["orig"] = np.arange(b, e)
                self.lib.sel_oracle_top1_se(cands.ctypes.data_as(C.c_void_p), e - b, C.byref(self.p), out[r: r + 1].ctypes.data_as(C.c_void_p))
        return out


def select_topn(selector: Selector, cand_begin: np.ndarray, scores: np.ndarray, topn: int):
    """ScoreBuffer::topNSE per read -> (sel int32 [n, topn] candidate indices in sorted order, -1 padded; n_sel, mapq, num_top int32 [n])."""
    n = len(cand_begin) - 1
    sel = np.full((n, topn), -1, np.int32)
    n_sel, mapq, num_top = np.zeros(n, np.int32), np.zeros(n, np.int32), np.ones(n, np.int32)
    for r in range(n):
        b, e = int(cand_begin[r]), int(cand_begin[r + 1])
        if e > b:
            cands = np.zeros(e - b, dtype=SEL_CAND)
            cands["score"] = scores[b:e]
            cands["orig"] = np.arange(b, e)
            row = np.zeros(max(topn, e - b), np.int32)
            ns, mq, nt = C.c_int(0), C.c_int(0), C.c_int(0)
            selector.lib.sel_oracle_topn_se(cands.ctypes.data_as(C.c_void_p), e - b, topn, C.byref(selector.p), row.ctypes.data_as(C.c_void_p), C.byref(ns),
                                            C.byref(mq), C.byref(nt))
            sel[r, : ns.value] = row[: ns.value]
            n_sel[r], mapq[r], num_top[r] = ns.value, mq.value, nt.value
    return sel, n_sel, mapq, num_top


def revcomp_row(row: np.ndarray) -> np.ndarray:
    s = row.tobytes().split(b"\0")[0]
    out = np.zeros_like(row)
    out[: len(s)] = np.frombuffer(s.translate(COMP)[::-1], np.uint8)
    return out


def _windows(packed, concat_len, reads, pairs, qml, cor, buf_len, fill_on_failure):
    """What ScoreBuffer (ScoreBuffer.cpp:92-118) / AlignmentBuffer (AlignmentBuffer.cpp:73-102) hand to BatchScore / BatchAlign."""
    n = len(pairs)
    refs = np.zeros((n, max(buf_len, qml + cor)), np.uint8)
    qrys = np.zeros((n, qml), np.uint8)
    for i, p in enumerate(pairs):
        w = port.decode_window(packed, concat_len, int(p["window_start"]), buf_len)
        if w is None:
            refs[i, :buf_len] = ord("N") if fill_on_failure else 0
        else:
            refs[i, :buf_len] = np.frombuffer(w, np.uint8)
        row = reads[int(p["read_index"])]
        qrys[i] = revcomp_row(row) if (int(p["flags"]) & 1) else row
    return refs, qrys


def map_batch_topn(packed: np.ndarray, concat_len: int, ix: cs_port.Index, reads: np.ndarray, qml: int, corridor: int, mode: int, sensitivity: float,
                   topn: int, selector: Optional[Selector] = None):
    """Single-end run with topn > 1 (ScoreBuffer::topNSE): up to topn alignments per read.  -> namespace with the fields
    nextgenmap_b200.host.pipeline.sam_lines_topn takes (sel [n, topn], n_sel, recs [n, topn], strings(r, j))."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n = reads.shape[0]
    begin, cands, max_hit = ix.search(reads, sensitivity)
    pairs = np.zeros(len(cands), dtype=PAIR)
    pairs["window_start"] = (cands["location"].astype(np.uint64) - np.uint64(corridor >> 1))
    pairs["read_index"] = np.repeat(np.arange(n, dtype=np.uint32), np.diff(begin))
    pairs["flags"] = np.where(cands["reverse"] != 0, 3, 0)
    refs, qrys = _windows(packed, concat_len, reads, pairs, qml, corridor, ((qml + corridor) | 1) + 1, True)
    scores = port.batch_score(refs, qrys, qml, corridor, mode) if len(pairs) else np.zeros(0, np.float32)
    selector = selector or Selector()
    sel, n_sel, mapq, num_top = select_topn(selector, begin, scores, topn)
    rr, jj = np.nonzero(sel >= 0)
    winners = pairs[sel[rr, jj]].copy()
    recs = np.zeros((n, topn), dtype=REC)
    recs["score"] = -1.0
    strings = {}
    if len(rr):
        refs, qrys = _windows(packed, concat_len, reads, winners, qml, corridor, (qml + corridor) | 2, False)
        for r, j, a in zip(rr, jj, port.batch_align(refs, qrys, qml, corridor, mode)):
            recs[r, j] = (a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore)
            strings[(int(r), int(j))] = (a.cigar, a.md_raw)
    return SimpleNamespace(cand_begin=begin, pairs=pairs, scores=scores, max_hit=max_hit, sel=sel, n_sel=n_sel, mapq=mapq, num_top=num_top, recs=recs,
                           strings=lambda r, j: strings[(int(r), int(j))])


def map_batch(packed: np.ndarray, concat_len: int, ix: cs_port.Index, reads: np.ndarray, qml: int, corridor: int, mode: int, sensitivity: float,
              selector: Optional[Selector] = None, paired: bool = False, mutate: Optional[dict] = None, scoring: Optional[port.Scoring] = None):
    """mutate: {"mode": 1 (--bs-mapping) | 2 (--slam-seq 4), "bs_cutoff": 6, "read_skip": 2} -> candidate search with CS::PrefixMutateSearch
    and the direction flag of ScoreBuffer / AlignmentBuffer (strand, inverted for second mates; ScoreBuffer.cpp:92-110, AlignmentBuffer.cpp:84-97)
    handed to BatchScore / BatchAlign; scoring: the run's scoring scheme (Config.cpp:463-469 under bs_mapping)."""
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n = reads.shape[0]
    sc = scoring or port.Scoring()
    if mutate:
        begin, cands, max_hit = ix.search_mut(reads, sensitivity, mutate["mode"], bs_cutoff=mutate.get("bs_cutoff", 6), paired=paired,
                                              read_skip=mutate.get("read_skip", 2 if mutate["mode"] == 1 else 0))
    else:
        begin, cands, max_hit = ix.search(reads, sensitivity)
    pairs = np.zeros(len(cands), dtype=PAIR)
    pairs["window_start"] = (cands["location"].astype(np.uint64) - np.uint64(corridor >> 1))
    pairs["read_index"] = np.repeat(np.arange(n, dtype=np.uint32), np.diff(begin))
    rev = cands["reverse"] != 0
    second = (pairs["read_index"] & 1).astype(bool) if (mutate and paired) else np.zeros(len(cands), bool)
    pairs["flags"] = np.where(rev, 1, 0) | np.where(rev ^ second, 2, 0)
    use_dirs = sc.bs_mapping == 1 or sc.slam_seq != 0           # ScoreBuffer.cpp:127: the direction buffer is only handed over in these modes
    dirs = ((pairs["flags"] >> 1) & 1).astype(np.uint8) if use_dirs else None
    refs, qrys = _windows(packed, concat_len, reads, pairs, qml, corridor, ((qml + corridor) | 1) + 1, True)
    scores = port.batch_score(refs, qrys, qml, corridor, mode, sc, dirs) if len(pairs) else np.zeros(0, np.float32)
    selector = selector or Selector()
    read_len = np.array([len(reads[r].tobytes().split(b"\0")[0]) for r in range(n)], np.int32)
    if paired:
        sel = selector.select_pairs(begin, cands["location"], scores, read_len)
    else:
        sel = selector.select_single(begin, scores)
    best = sel["best"].astype(np.int32)
    has = np.nonzero(best >= 0)[0]
    winners = pairs[best[has]].copy()
    winners["read_index"] = has
    recs = np.zeros(n, dtype=REC)
    recs["score"] = -1.0
    strings = {}
    if len(has):
        refs, qrys = _windows(packed, concat_len, reads, winners, qml, corridor, (qml + corridor) | 2, False)
        wdirs = ((winners["flags"] >> 1) & 1).astype(np.uint8) if use_dirs else None
        for r, a in zip(has, port.batch_align(refs, qrys, qml, corridor, mode, sc, wdirs)):
            recs[r] = (a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore)
            strings[int(r)] = (a.cigar, a.md_raw)
    return SimpleNamespace(cand_begin=begin, pairs=pairs, scores=scores, max_hit=max_hit, best_pair=best, mapq=sel["mapq"].astype(np.int32),
                           num_top=sel["num_top"].astype(np.int32), recs=recs, heap=None, strings=lambda r: strings[int(r)],
                           pair_fail=sel["paired_fail"].astype(np.int32) if paired else None, insert=sel["insert"].astype(np.int32))
