"""Host-side mirror of NextGenMap's single-end mapping run on top of the device library:

    reads -> CS (k-mer vote, CS.cpp) -> ScoreBuffer (BatchScore of every candidate, top1SE + computeMQ)
          -> AlignmentBuffer (BatchAlign of the winner) -> GenericReadWriter filters -> SAMWriter lines

Everything numeric runs in ``libngm_b200.so`` (there is no CPU path); this module only strings the device calls together
and renders the records the way ``SAMWriter::DoWriteReadGeneric`` / ``DoWriteUnmappedReadGeneric`` do
(src/writer/SAMWriter.cpp:98-228,312-365), after the post-processing of ``AlignmentBuffer::DoRun`` / ``WriteRead``
(src/AlignmentBuffer.cpp:121-135,166-176) and the output filters of ``GenericReadWriter::WriteRead``
(src/writer/GenericReadWriter.h:190-256).  Single-end, ``topn`` 1, no bs-mapping -- what BASELINE configs[0]/[1] run.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .cuda_sw import ALIGN_REC, PAIR, CudaSW

COMP = bytes.maketrans(b"ACGT", b"TGCA")


@dataclass
class MappedBatch:
    """Per-read results of one batch (arrays of length n_reads unless noted)."""
    cand_begin: np.ndarray      # int32 [n + 1]
    pairs: np.ndarray           # PAIR [total]  candidates in the reference's order
    scores: np.ndarray          # float32 [total]  BatchScore
    max_hit: np.ndarray         # float32  MappedRead::s
    best_pair: np.ndarray       # int32  index into pairs, -1 = no candidate
    mapq: np.ndarray            # int32
    num_top: np.ndarray         # int32  MappedRead::numTopScores
    recs: np.ndarray            # ALIGN_REC
    heap: np.ndarray            # uint8 string heap (CIGAR, MD)


def map_reads(sw: CudaSW, reads: np.ndarray, mode: int = 0) -> MappedBatch:
    """One batch through candidate search, scoring, top-1 selection and alignment.  Needs set_reference + a prefix table."""
    import torch
    reads = np.ascontiguousarray(reads, dtype=np.uint8)
    n = reads.shape[0]
    begin, pairs, votes, max_hit = sw.cs_search(reads)
    sw.set_reads(reads)
    scores = sw.score_pairs(mode, pairs)
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
    winners = np.zeros(n, dtype=PAIR)
    has = best >= 0
    winners[has] = pairs[best[has]]
    winners["read_index"] = np.arange(n, dtype=np.uint32)
    winners["flags"][~has] = 4                     # NGM_B200_PAIR_SKIP
    recs, heap = sw.align_pairs(mode, winners)
    return MappedBatch(begin, pairs, scores, max_hit, best, d_mq.cpu().numpy(), d_nt.cpu().numpy(), recs, heap)


def _xi(identity: float) -> str:
    # SAMWriter.cpp:188-189: round(Identity * 10000.0f) / 10000.0f, printed with %g
    v = float(np.float32(identity) * np.float32(10000.0))
    r = math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)
    return "%g" % float(np.float32(r / 10000.0))


def sam_lines(sw: CudaSW, batch: MappedBatch, reads: np.ndarray, names: Sequence[str], quals: Sequence[bytes], encref, corridor: int,
              min_identity: float = 0.65, min_residues: float = 0.5) -> List[str]:
    """SAM body lines of a single-end batch (no header), one per read, in read order."""
    out = []
    for r in range(reads.shape[0]):
        seq = reads[r].tobytes().split(b"\0")[0]
        length = len(seq)
        qual = quals[r]
        name = names[r]
        line = None
        bp = int(batch.best_pair[r])
        rec = batch.recs[r]
        if bp >= 0 and float(rec["score"]) >= 0.0:
            p = batch.pairs[bp]
            reverse = bool(int(p["flags"]) & 1)
            # AlignmentBuffer.cpp:129: Location += PositionOffset - (corridor >> 1); window_start is Location - corridor/2 already
            loc = (int(p["window_start"]) + int(rec["position_offset"])) & (2 ** 64 - 1)
            conv = encref.convert(loc)              # AlignmentBuffer.cpp:171-175
            qstart, qend = int(rec["qstart"]), int(rec["qend"])
            mres = length * min_residues if min_residues <= 1.0 else min_residues      # GenericReadWriter.h:206-210
            mapped = conv is not None and float(rec["identity"]) >= np.float32(min_identity) and float(length - qstart - qend) >= np.float32(mres)
            if mapped:
                contig, pos = conv
                cigar, md = sw.strings_of(batch.recs, batch.heap, r)
                s, q = seq, qual
                flags = 0
                if reverse:
                    s = seq.translate(COMP)[::-1]
                    q = qual[::-1]
                    flags |= 0x10
                ntop = int(batch.num_top[r])
                line = "\t".join([name, str(flags), encref.contigs[contig][0], str((pos + 1) & 0xFFFFFFFF), str(int(batch.mapq[r])), cigar.decode(), "*", "0", "0",
                                  s.decode(), q.decode(), "AS:i:%d" % int(batch.scores[bp]), "NM:i:%d" % int(rec["nm"]), "NH:i:%d" % ntop,
                                  "XI:f:" + _xi(float(rec["identity"])), "X0:i:%d" % ntop, "XE:i:%d" % int(batch.max_hit[r]),
                                  "XR:i:%d" % (length - qstart - qend), "MD:Z:" + md.split(b"\0")[0].decode()])
        if line is None:                            # SAMWriter.cpp:312-365
            line = "\t".join([name, "4", "*", "0", "0", "*", "*", "0", "0", seq.decode(), qual.decode()])
        out.append(line)
    return out
