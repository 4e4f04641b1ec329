"""ngm_b200_run_batch: ScoreBuffer::DoRun + AlignmentBuffer::DoRun for a whole batch behind one pipelined C-ABI call.

Checked three ways:
* against the oracle (window decode the way ScoreBuffer / AlignmentBuffer do it + BatchScore / BatchAlign restatements + the top1SE rule),
  which pins the fused path of single-candidate reads (their score comes out of the alignment's forward pass);
* every input format (ASCII / 2-bit packed reads, 16-byte / 64-bit descriptors), several lanes and sub-batch sizes, `strata`
  must give identical outputs;
* paired batches: the running insert-size sums are one sequence over all lanes -- sub-batching must not change any result.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import port
from tests import util
from tests.test_gpu_descriptor_path import build_case, host_windows

pytestmark = pytest.mark.gpu


def top1_host(begin, scores, strata=False):
    """ScoreBuffer::top1SE + computeMQ (ScoreBuffer.cpp:34-40,228-277)."""
    n = len(begin) - 1
    best, mapq, ntop = np.full(n, -1, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    for r in range(n):
        b, s, bi, nb = np.float32(0), np.float32(0), 0, 0
        for j in range(begin[r], begin[r + 1]):
            v = scores[j]
            if v > s:
                if v > b:
                    s, b, bi, nb = b, v, j - begin[r], 1
                elif v == b:
                    nb += 1
                    s = b
                else:
                    s = v
            elif v == b:
                nb += 1
        mq = int(np.ceil(np.float32(60.0) * np.float32(b - s) / np.float32(b))) if (b > 0 and s >= 0) else 0
        if begin[r + 1] > begin[r]:
            best[r] = begin[r] + bi
            if strata and nb != 1:
                best[r], mq, nb = -1, 0, 1
        mapq[r], ntop[r] = mq, nb
    return best, mapq, ntop


def make_batch(seed, n_reads, qml, cor):
    concat, packed, reads, pairs = build_case(seed, n_reads, qml, cor)
    rng = np.random.default_rng(seed + 1)
    # reads without candidates, reads with N / IUPAC bases, an empty read, reads with many (equal) candidates
    drop = set(int(r) for r in rng.choice(n_reads, n_reads // 25, replace=False))
    for r in rng.choice(n_reads, n_reads // 15, replace=False):
        L = int(np.count_nonzero(reads[r]))
        if L > 10:
            reads[r, int(rng.integers(0, L))] = ord("N")
            if rng.random() < 0.3:
                reads[r, int(rng.integers(0, L))] = ord("R")
    reads[7] = 0
    keep = [p for p in pairs if int(p["read_index"]) not in drop]
    extra = []
    for r in rng.choice(n_reads, n_reads // 40, replace=False):
        if int(r) in drop:
            continue
        mine = [p for p in keep if int(p["read_index"]) == int(r)]
        if mine:
            extra += [mine[0]] * int(rng.integers(1, 3))                   # the same window again: equal best scores (strata)
            for _ in range(int(rng.integers(0, 4))):
                q = mine[0].copy()
                q["window_start"] = int(rng.integers(0, len(concat)))
                extra.append(q)
    allp = np.array(keep + extra, dtype=pairs.dtype)
    order = np.argsort(allp["read_index"], kind="stable")
    allp = allp[order]
    begin = np.searchsorted(allp["read_index"], np.arange(n_reads + 1)).astype(np.int32)
    return concat, packed, reads, allp, begin


def expected_from_oracle(concat, packed, reads, pairs, begin, qml, cor, mode, strata):
    score_buf = ((qml + cor) | 1) + 1            # ScoreBuffer.h:112
    align_buf = (qml + cor) | 2                  # AlignmentBuffer.h:67
    refs, qrys = host_windows(packed, len(concat), reads, pairs, qml, cor, score_buf, True)
    # the oracle models the CPU device's quad rule (an empty read in lane 0 silences its whole quad, oclSwScore.cl:124); on the
    # descriptor path an empty read is its own quad leader: keep it out of the oracle's quads and put its result in by hand
    empty = np.array([not reads[int(p["read_index"])].any() for p in pairs])
    qrys[empty, 0] = ord("A")
    scores = port.batch_score(refs, qrys, qml, cor, mode)
    scores[empty] = -1.0 if mode == 0 else -16000.0
    best, mapq, ntop = top1_host(begin, scores, strata)
    win = pairs[best[best >= 0]]
    refs, qrys = host_windows(packed, len(concat), reads, win, qml, cor, align_buf, False)
    qrys[~qrys.any(axis=1), 0] = ord("A")
    aligns = port.batch_align(refs, qrys, qml, cor, mode)
    return scores, best, mapq, ntop, aligns


def undefined_pairs(pairs, concat_len, qml, cor):
    """odd window offsets that run past the end of the reference decode one nibble of uninitialised memory in the reference
    (SequenceProvider.cpp:417-423); windows whose decode fails keep a stale buffer in AlignmentBuffer (AlignmentBuffer.cpp:101)"""
    starts = pairs["window_start"].astype(np.uint64)
    score_buf = ((qml + cor) | 1) + 1
    odd_end = ((starts & np.uint64(1)) == 1) & (starts < concat_len) & (starts + np.uint64(score_buf - 2) >= concat_len)
    return odd_end, starts >= concat_len


def rec_tuple(sw, recs, heap, i):
    r = recs[i]
    cig, md = sw.strings_of(recs, heap, i)
    return util.align_tuple(r["position_offset"], r["qstart"], r["qend"], r["nm"], r["identity"], r["score"], cig, md.split(b"\0")[0])


@pytest.mark.parametrize("qml,cor", [(152, 27), (102, 20), (252, 42), (252, 80)])
@pytest.mark.parametrize("mode", [0, 1])
def test_run_batch_matches_oracle(qml, cor, mode):
    from nextgenmap_b200.host import CudaSW
    n_reads = 1800
    concat, packed, reads, pairs, begin = make_batch(4242 + qml + cor, n_reads, qml, cor)
    sw = CudaSW(qml, cor)
    sw.set_reference(packed, len(concat))
    sw.set_pipeline(3, 512)
    scores, best, mapq, ntop, aligns = expected_from_oracle(concat, packed, reads, pairs, begin, qml, cor, mode, False)
    odd_end, failed = undefined_pairs(pairs, len(concat), qml, cor)
    got = sw.run_batch(mode, reads, begin, pairs)
    ok_pair = ~odd_end
    np.testing.assert_array_equal(util.bits(got["scores"][ok_pair]), util.bits(scores[ok_pair]))
    # selection is compared where every candidate of the read is defined
    read_ok = np.array([ok_pair[begin[r]:begin[r + 1]].all() for r in range(n_reads)])
    np.testing.assert_array_equal(got["best_pair"][read_ok], best[read_ok])
    np.testing.assert_array_equal(got["mapq"][read_ok], mapq[read_ok])
    np.testing.assert_array_equal(got["num_top"][read_ok], ntop[read_ok])
    bad, k = [], 0
    for r in range(n_reads):
        if best[r] < 0:
            assert got["recs"][r]["score"] == -1.0
            continue
        a = aligns[k]
        k += 1
        if not read_ok[r] or failed[best[r]]:
            continue
        if not reads[r].any():                                             # empty read: the lane is skipped, failure convention
            assert got["recs"][r]["score"] == -1.0
            continue
        if a.ascore == -1.0 and a.cigar == b"!!!":
            g, w = (int(got["recs"][r]["position_offset"]), float(got["recs"][r]["score"])), (a.position_offset, -1.0)
        else:
            g = rec_tuple(sw, got["recs"], got["heap"], r)
            w = util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md)
        if g != w:
            bad.append((r, g, w))
    assert not bad, f"{len(bad)} alignments differ, first {bad[0]}"
    assert sw.launch_count() > 0
    sw.close()


def same_results(sw, a, b, n_reads):
    for f in ("scores", "best_pair", "mapq", "num_top"):
        np.testing.assert_array_equal(util.bits(a[f]) if f == "scores" else a[f], util.bits(b[f]) if f == "scores" else b[f], err_msg=f)
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        np.testing.assert_array_equal(a["recs"][f], b["recs"][f], err_msg=f)
    np.testing.assert_array_equal(util.bits(a["recs"]["identity"]), util.bits(b["recs"]["identity"]))
    for r in range(n_reads):
        if a["recs"][r]["score"] >= 0:
            assert sw.strings_of(a["recs"], a["heap"], r) == sw.strings_of(b["recs"], b["heap"], r), r


@pytest.mark.parametrize("strata", [0, 1])
def test_run_batch_formats_lanes_and_strata(strata):
    from nextgenmap_b200.host import CudaSW
    qml, cor, n_reads = 152, 27, 3000
    concat, packed, reads, pairs, begin = make_batch(99 + strata, n_reads, qml, cor)
    sw = CudaSW(qml, cor)
    sw.set_reference(packed, len(concat))
    sw.se_configure(strata)
    sw.set_pipeline(3, 700)
    base = sw.run_batch(0, reads, begin, pairs)
    # strata against the host rule on the batch's own scores
    best, mapq, ntop = top1_host(begin, base["scores"], bool(strata))
    np.testing.assert_array_equal(base["best_pair"], best)
    np.testing.assert_array_equal(base["mapq"], mapq)
    np.testing.assert_array_equal(base["num_top"], ntop)
    if strata:
        assert (best[(begin[1:] - begin[:-1]) > 0] < 0).any(), "the case must hold reads with several equally best candidates"
    for packed_reads, u64, lanes, sb in [(True, False, 3, 700), (False, True, 2, 1024), (True, True, 1, 4096), (True, True, 4, 256)]:
        sw2 = CudaSW(qml, cor)
        sw2.set_reference(packed, len(concat))
        sw2.se_configure(strata)
        sw2.set_pipeline(lanes, sb)
        got = sw2.run_batch(0, reads, begin, pairs, packed=packed_reads, desc_u64=u64)
        same_results(sw, base, got, n_reads)
        sw2.close()
    # a heap that is too small is reported with a sufficient size, and the repeat succeeds (run_batch retries inside the mirror)
    small = sw.run_batch(0, reads, begin, pairs, str_capacity=64)
    same_results(sw, base, small, n_reads)
    # the classic device entry point honours strata as well
    import torch
    d_begin, d_scores = torch.from_numpy(begin).cuda(), torch.from_numpy(base["scores"]).cuda()
    d_best, d_mq, d_nt = (torch.empty(n_reads, dtype=torch.int32, device="cuda") for _ in range(3))
    st = torch.cuda.current_stream().cuda_stream
    assert sw.lib.ngm_b200_dev_select_top1_ex(sw.ctx, n_reads, d_begin.data_ptr(), d_scores.data_ptr(), d_best.data_ptr(), d_mq.data_ptr(), d_nt.data_ptr(), st) == n_reads
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d_best.cpu().numpy(), best)
    np.testing.assert_array_equal(d_mq.cpu().numpy(), mapq)
    sw.close()


def test_run_batch_paired_is_independent_of_sub_batching():
    from nextgenmap_b200.host import CudaSW
    from oracle import mapper_port
    qml, cor, n_reads = 102, 20, 2400
    concat, packed, reads, pairs, begin = make_batch(777, n_reads, qml, cor)
    # mates: make the candidates of rows 2f / 2f + 1 lie close to each other for most fragments
    rng = np.random.default_rng(3)
    for f in range(n_reads // 2):
        a0, a1, b0, b1 = begin[2 * f], begin[2 * f + 1], begin[2 * f + 1], begin[2 * f + 2]
        if a1 > a0 and b1 > b0 and rng.random() < 0.8:
            pairs["window_start"][b0] = (int(pairs["window_start"][a0]) + int(rng.integers(150, 500))) % len(concat)
    out = []
    for lanes, sb in [(1, 1 << 20), (3, 400), (2, 1000)]:
        sw = CudaSW(qml, cor)
        sw.set_reference(packed, len(concat))
        sw.pe_configure()
        sw.set_pipeline(lanes, sb)
        out.append((sw.run_batch(0, reads, begin, pairs, paired=True, packed=True, desc_u64=True), sw.pe_insert_stats(), sw))
    base, stats0, sw0 = out[0]
    for got, stats, _ in out[1:]:
        same_results(sw0, base, got, n_reads)
        np.testing.assert_array_equal(base["pair_fail"], got["pair_fail"])
        assert stats == stats0
    # and the selection equals the oracle's (top1PE restatement) on the batch's own scores
    lens = np.array([int(np.count_nonzero(r)) for r in reads], np.int32)
    want = mapper_port.Selector().select_pairs(begin, pairs["window_start"] + np.uint64(cor >> 1), base["scores"], lens)
    np.testing.assert_array_equal(base["best_pair"], want["best"])
    np.testing.assert_array_equal(base["mapq"], want["mapq"])
    np.testing.assert_array_equal(base["pair_fail"], want["paired_fail"])
    # the records are those of the classic align entry point for every winner -- including the winners of reads with more than four candidates,
    # which the batch aligns in a second pass over a late list
    counts = begin[1:] - begin[:-1]
    has = base["best_pair"] >= 0
    assert np.count_nonzero(has & (counts > 4)) >= 5 and np.count_nonzero(has & (counts <= 4)) > 1000
    sw0.set_reads(reads)
    winners = pairs[base["best_pair"][has]].copy()
    winners["read_index"] = np.nonzero(has)[0]
    recs, heap = sw0.align_pairs(0, winners)
    got = base["recs"][has]
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        np.testing.assert_array_equal(got[f], recs[f], err_msg=f)
    rows = np.nonzero(has)[0]
    for i in range(len(rows)):
        if recs[i]["score"] >= 0:
            assert sw0.strings_of(base["recs"], base["heap"], int(rows[i])) == sw0.strings_of(recs, heap, i), i
    assert np.all(base["recs"]["score"][~has] == -1.0)
    for _, _, sw in out:
        sw.close()


def test_first_generation_forward_kernel_still_agrees():
    """NGM_B200_FWD=1 runs the first-generation forward kernel (register snapshot); both generations must give the same batch."""
    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from tests.test_gpu_batch import make_batch
from nextgenmap_b200.host import CudaSW
concat, packed, reads, pairs, begin = make_batch(31, 2000, 152, 27)
sw = CudaSW(152, 27); sw.set_reference(packed, len(concat))
g = sw.run_batch(0, reads, begin, pairs)
np.savez(sys.argv[1], scores=g["scores"], best=g["best_pair"], recs=g["recs"].view(np.uint8), used=g["str_used"])
''' % (str(util.ROOT),)
    import tempfile
    res = []
    with tempfile.TemporaryDirectory() as td:
        for v in ("0", "1"):
            env = dict(os.environ, NGM_B200_FWD=v)
            path = os.path.join(td, f"o{v}.npz")
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, cwd=str(util.ROOT))
            res.append(dict(np.load(path)))
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC
    a, b = (r["recs"].view(ALIGN_REC).reshape(-1) for r in res)
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        np.testing.assert_array_equal(a[f], b[f], err_msg=f)
    np.testing.assert_array_equal(res[0]["scores"], res[1]["scores"])
    assert int(res[0]["used"]) == int(res[1]["used"])


@pytest.mark.parametrize("qml,cor,paired,switch", [(152, 27, False, "NGM_B200_FWD_EXACT"), (152, 27, False, "NGM_B200_FWD_ALL"), (152, 27, True, "NGM_B200_PE_FWD_ALL"),
                                                  (252, 80, False, "NGM_B200_WIDE_V2"), (252, 80, True, "NGM_B200_WIDE_V2"), (152, 27, False, "NGM_B200_STAGGER")])
def test_routes_agree(qml, cor, paired, switch):
    """Every A/B switch of the batch engine selects another route to the SAME results: exact-corridor kernels, forward pass over every
    candidate (single-end and paired), the second-generation kernel on wide bands, staggered lanes.  <switch>=0 against the default."""
    code = r'''
import numpy as np, sys
sys.path.insert(0, %r)
from tests.test_gpu_batch import make_batch
from nextgenmap_b200.host import CudaSW
qml, cor, paired = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4] == "1"
concat, packed, reads, pairs, begin = make_batch(57, 1600, qml, cor)
sw = CudaSW(qml, cor); sw.set_reference(packed, len(concat))
if paired:
    sw.pe_configure()
sw.set_pipeline(3, 300)
g = sw.run_batch(0, reads, begin, pairs, paired=paired, packed=True, desc_u64=True)
strings = [sw.strings_of(g["recs"], g["heap"], r) for r in range(len(reads)) if g["recs"][r]["score"] >= 0]
np.savez(sys.argv[1], scores=g["scores"], best=g["best_pair"], mapq=g["mapq"], ntop=g["num_top"], pf=g["pair_fail"], recs=g["recs"].view(np.uint8),
         strings=np.frombuffer(b"\n".join(c + b"\t" + m for c, m in strings), np.uint8))
''' % (str(util.ROOT),)
    import tempfile
    res = []
    with tempfile.TemporaryDirectory() as td:
        for v in (None, "0"):
            env = dict(os.environ)
            env.pop(switch, None)
            if v is not None:
                env[switch] = v
            path = os.path.join(td, f"o{v}.npz")
            subprocess.run([sys.executable, "-c", code, path, str(qml), str(cor), "1" if paired else "0"], check=True, env=env, cwd=str(util.ROOT))
            res.append(dict(np.load(path)))
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC
    a, b = (r["recs"].view(ALIGN_REC).reshape(-1) for r in res)
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        np.testing.assert_array_equal(a[f], b[f], err_msg=f)
    for f in ("scores", "best", "mapq", "ntop", "strings") + (("pf",) if paired else ()):
        np.testing.assert_array_equal(res[0][f], res[1][f], err_msg=f)
    assert np.count_nonzero(a["score"] >= 0) > 1000
