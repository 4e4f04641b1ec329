"""Descriptor fast path: reference resident in HBM in NGM's own 4-bit packing, reads uploaded once,
(read, window_start, strand) descriptors instead of host-decoded windows.

Oracle: the same pairs pushed through the oracle's restatement of what NGM's callers do on the host --
DecodeRefSequence into a window buffer (ScoreBuffer.cpp:113-118 with refMaxLen of ScoreBuffer.h:112,
AlignmentBuffer.cpp:101-102 with refMaxLen of AlignmentBuffer.h:67), reverse complement for minus-strand
candidates (MappedRead.cpp:53-67) -- followed by BatchScore / BatchAlign.
"""
import numpy as np
import pytest

from oracle import port
from tests import util

pytestmark = pytest.mark.gpu

COMP = {ord("A"): ord("T"), ord("T"): ord("A"), ord("C"): ord("G"), ord("G"): ord("C")}


def revcomp_row(row: np.ndarray) -> np.ndarray:
    s = row.tobytes().split(b"\0")[0]
    out = np.zeros_like(row)
    rc = bytes(COMP.get(b, b) for b in reversed(s))
    out[: len(rc)] = np.frombuffer(rc, np.uint8)
    return out


def build_case(seed, n_reads, qml, cor, ref_len=200_000):
    rng = np.random.default_rng(seed)
    L = qml - 2 if qml % 2 == 0 else qml - 1
    # concatenated reference the way SequenceProvider lays it out: 1000 N, contig, 1000 N, contig, 1000 N
    contig = lambda n: bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)])
    c1, c2 = contig(ref_len // 2), contig(ref_len // 2 - 37)
    c1 = c1[:5000] + b"N" * 23 + c1[5023:]
    concat = b"N" * 1000 + c1 + b"N" * 1000 + c2 + (b"N" if len(c2) & 1 else b"") + b"N" * 1000
    packed = port.pack_ref(concat)
    reads = np.zeros((n_reads, qml), np.uint8)
    pairs = []
    cb = np.frombuffer(concat, np.uint8)
    for r in range(n_reads):
        kind = r % 10
        length = L if kind != 3 else int(rng.integers(20, L))
        pos = int(rng.integers(900, len(concat) - length - 50))
        if kind == 4:
            pos = int(rng.integers(len(concat) - length - 5, len(concat) - 10))       # window runs past the end -> 'x'
        if kind == 5:
            pos = int(rng.integers(0, cor))                                             # loc - corridor/2 underflows
        seq = cb[pos: pos + length].copy()
        seq = seq[seq != 0][:length]
        mut = rng.random(len(seq)) < 0.03
        seq[mut] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(mut.sum()))]
        if kind == 6 and len(seq) > 40:                                                 # small deletion in the read
            seq = np.concatenate([seq[:30], seq[33:]])
        if kind == 7 and len(seq) > 40:                                                 # small insertion
            seq = np.concatenate([seq[:30], np.frombuffer(b"GA", np.uint8), seq[30:]])[:length]
        reverse = bool(rng.integers(0, 2))
        row = np.zeros(qml, np.uint8)
        row[: len(seq)] = seq
        if reverse:
            row = revcomp_row(row)          # the read as sequenced; RevSeq gives back `seq`
        reads[r] = row
        jitter = int(rng.integers(-3, 4))
        start = (pos + jitter - (cor >> 1)) % (1 << 64)
        pairs.append((start, r, 1 if reverse else 0))
        if kind in (1, 2):                  # a decoy candidate somewhere else, other strand
            pairs.append((int(rng.integers(0, len(concat))), r, 0 if reverse else 1))
    parr = np.zeros(len(pairs), dtype=[("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])
    for i, (s, r, f) in enumerate(pairs):
        parr[i] = (s, r, f)
    return concat, packed, reads, parr


def host_windows(packed, concat_len, reads, pairs, qml, cor, buf_len, fill_on_failure):
    """What ScoreBuffer / AlignmentBuffer hand to BatchScore / BatchAlign."""
    n = len(pairs)
    refs = np.zeros((n, max(buf_len, qml + cor)), np.uint8)
    qrys = np.zeros((n, qml), np.uint8)
    for i, p in enumerate(pairs):
        w = port.decode_window(packed, concat_len, int(p["window_start"]), buf_len)
        if w is None:
            refs[i, :buf_len] = ord("N") if fill_on_failure else 0      # ScoreBuffer.cpp:117
        else:
            refs[i, :buf_len] = np.frombuffer(w, np.uint8)
        row = reads[int(p["read_index"])]
        qrys[i] = revcomp_row(row) if (int(p["flags"]) & 1) else row
    return refs, qrys


@pytest.mark.parametrize("qml,cor", [(152, 27), (102, 20), (252, 42)])
@pytest.mark.parametrize("lane_mode", [0, 1])
def test_descriptor_scores_and_alignments(qml, cor, lane_mode):
    from nextgenmap_b200.host import CudaSW
    concat, packed, reads, pairs = build_case(1234 + qml, 1500, qml, cor)
    sw = CudaSW(qml, cor, lane_mode=lane_mode)
    sw.set_reference(packed, len(concat))
    sw.set_reads(reads)
    score_buf = ((qml + cor) | 1) + 1            # ScoreBuffer.h:112
    align_buf = (qml + cor) | 2                  # AlignmentBuffer.h:67 (operator precedence: | 1 + 1)
    # For an odd offset DecodeRefSequence decodes one nibble more than asked for
    # (SequenceProvider.cpp:417-423); when the window also runs past the end of the concatenated
    # reference that nibble lies beyond the encoded array (uninitialised memory in the reference).
    # The device defines every position >= concat_len as 'x'; such pairs are left out of the comparison.
    starts = pairs["window_start"].astype(np.uint64)
    undefined = ((starts & np.uint64(1)) == 1) & (starts < len(concat)) & (starts + np.uint64(score_buf - 2) >= len(concat))
    for mode in (0, 1):
        refs, qrys = host_windows(packed, len(concat), reads, pairs, qml, cor, score_buf, True)
        want = port.batch_score(refs, qrys, qml, cor, mode)
        got = sw.score_pairs(mode, pairs)
        # the oracle models the CPU device's quad-granular empty-read rule; no read is empty here
        np.testing.assert_array_equal(util.bits(got[~undefined]), util.bits(want[~undefined]), err_msg=f"mode {mode}")
        # alignments: windows as AlignmentBuffer decodes them; pairs whose decode fails keep stale
        # buffers in the reference (AlignmentBuffer.cpp:101 ignores the return value) -> excluded
        ok = np.array([int(p["window_start"]) < len(concat) for p in pairs]) & ~undefined
        refs, qrys = host_windows(packed, len(concat), reads, pairs, qml, cor, align_buf, False)
        wa = port.batch_align(refs[ok], qrys[ok], qml, cor, mode)
        recs, heap = sw.align_pairs(mode, pairs[ok])
        bad = []
        for i, a in enumerate(wa):
            cig, md = sw.strings_of(recs, heap, i)
            r = recs[i]
            if a.ascore == -1.0 and a.cigar == b"!!!":
                g = (int(r["position_offset"]), float(r["score"]))
                w = (a.position_offset, -1.0)
            else:
                g = util.align_tuple(r["position_offset"], r["qstart"], r["qend"], r["nm"], r["identity"], r["score"], cig, md.split(b"\0")[0])
                w = util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md)
            if g != w:
                bad.append((i, g, w))
        assert not bad, f"mode {mode}: {len(bad)} differ, first {bad[0]}"
    sw.close()


def test_select_top1_matches_host_rule():
    """ScoreBuffer::top1SE + computeMQ (ScoreBuffer.cpp:34-40,228-277) on device."""
    import torch
    from nextgenmap_b200.host import CudaSW
    rng = np.random.default_rng(5)
    n_reads = 5000
    counts = rng.integers(0, 6, n_reads)
    begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    scores = rng.choice([-1.0, 0.0, 50.0, 100.0, 250.0, 300.0, 1480.0, 1500.0], int(begin[-1])).astype(np.float32)
    sw = CudaSW(152, 27)
    d_begin, d_scores = torch.from_numpy(begin).cuda(), torch.from_numpy(scores).cuda()
    d_best, d_mq = torch.empty(n_reads, dtype=torch.int32, device="cuda"), torch.empty(n_reads, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    assert sw.lib.ngm_b200_dev_select_top1(sw.ctx, n_reads, d_begin.data_ptr(), d_scores.data_ptr(), d_best.data_ptr(), d_mq.data_ptr(), st) == n_reads
    torch.cuda.synchronize()
    best, mq = d_best.cpu().numpy(), d_mq.cpu().numpy()
    for r in range(n_reads):
        b, s, bi = 0.0, 0.0, 0
        for j in range(begin[r], begin[r + 1]):
            v = float(scores[j])
            if v > s:
                if v > b:
                    s, b, bi = b, v, j - begin[r]
                elif v == b:
                    s = b
                else:
                    s = v
        want_mq = int(np.ceil(np.float32(60.0) * np.float32(b - s) / np.float32(b))) if (b > 0 and s >= 0) else 0
        want_best = begin[r] + bi if counts[r] else -1
        assert (best[r], mq[r]) == (want_best, want_mq), r
    sw.close()


@pytest.mark.parametrize("qml,cor", [(152, 27), (76, 16), (252, 42), (402, 65)])
def test_scored_align_equals_plain_align(qml, cor):
    """Resident pipeline: score -> top1 -> gather_winners_scored -> align_pairs_scored (the winners' known local maxima
    steer the forward pass) must give the records and strings of the plain align_pairs call."""
    import torch
    from nextgenmap_b200.host import CudaSW
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC
    concat, packed, reads, pairs = build_case(99 + qml, 3000, qml, cor)
    sw = CudaSW(qml, cor)
    sw.set_reference(packed, len(concat))
    sw.set_reads(reads)
    n_reads = len(reads)
    begin = np.searchsorted(pairs["read_index"], np.arange(n_reads + 1)).astype(np.int32)
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    d_pairs = torch.from_numpy(pairs.view(np.uint8).reshape(-1, 16)).to(dev)
    d_begin = torch.from_numpy(begin).to(dev)
    d_scores = torch.empty(len(pairs), dtype=torch.float32, device=dev)
    d_best = torch.empty(n_reads, dtype=torch.int32, device=dev)
    d_mq = torch.empty(n_reads, dtype=torch.int32, device=dev)
    d_wp = torch.empty((n_reads, 16), dtype=torch.uint8, device=dev)
    d_ws = torch.empty(n_reads, dtype=torch.float32, device=dev)
    lib, ctx = sw.lib, sw.ctx
    assert lib.ngm_b200_dev_score_pairs(ctx, 0, len(pairs), d_pairs.data_ptr(), d_scores.data_ptr(), st) == len(pairs)
    assert lib.ngm_b200_dev_select_top1(ctx, n_reads, d_begin.data_ptr(), d_scores.data_ptr(), d_best.data_ptr(), d_mq.data_ptr(), st) == n_reads
    assert lib.ngm_b200_dev_gather_winners_scored(ctx, n_reads, d_pairs.data_ptr(), d_scores.data_ptr(), d_best.data_ptr(), d_wp.data_ptr(),
                                                  d_ws.data_ptr(), st) == n_reads
    out = []
    for scored in (False, True):
        d_recs = torch.zeros((n_reads, 32), dtype=torch.uint8, device=dev)
        cap = 256 * n_reads
        d_str = torch.zeros(cap, dtype=torch.uint8, device=dev)
        d_cur = torch.zeros(1, dtype=torch.int32, device=dev)
        if scored:
            rc = lib.ngm_b200_dev_align_pairs_scored(ctx, 0, n_reads, d_wp.data_ptr(), d_ws.data_ptr(), d_recs.data_ptr(), d_str.data_ptr(), cap,
                                                     d_cur.data_ptr(), st)
        else:
            rc = lib.ngm_b200_dev_align_pairs(ctx, 0, n_reads, d_wp.data_ptr(), d_recs.data_ptr(), d_str.data_ptr(), cap, d_cur.data_ptr(), st)
        assert rc == n_reads, sw._err()
        torch.cuda.synchronize()
        assert int(d_cur.item()) <= cap
        recs = d_recs.cpu().numpy().view(ALIGN_REC).reshape(-1)
        heap = d_str.cpu().numpy()
        out.append([(int(r["position_offset"]), int(r["qstart"]), int(r["qend"]), int(r["nm"]), util.bits(np.array([r["identity"]]))[0],
                     float(r["score"])) + sw.strings_of(recs, heap, i) for i, r in enumerate(recs)])
    ws = d_ws.cpu().numpy()
    best = d_best.cpu().numpy()
    sc = d_scores.cpu().numpy()
    assert np.array_equal(ws[best >= 0], sc[best[best >= 0]])
    bad = [(i, a, b) for i, (a, b) in enumerate(zip(*out)) if a != b]
    assert not bad, f"{len(bad)} differ, first {bad[0]}"
    assert sum(1 for a in out[0] if a[5] >= 0) > n_reads // 2
    sw.close()
