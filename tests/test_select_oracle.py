"""oracle/select_oracle.c by itself: its step-by-step restatement of libstdc++'s std::sort leaves equal scores in the order std::sort does
(that order decides which candidate ScoreBuffer::top1PE / topNSE pick), and the small selection rules behave as ScoreBuffer.cpp states.
The whole-run pins against the unmodified NextGenMap are in tests/test_mapper_oracle.py."""
import subprocess
from pathlib import Path

import numpy as np

from oracle import mapper_port, port

ROOT = Path(__file__).resolve().parents[1]


def test_sort_restatement_equals_std_sort(tmp_path):
    port.build()
    exe = tmp_path / "sort_check"
    subprocess.run(["gcc", "-O2", "-c", "-o", str(tmp_path / "sel.o"), str(ROOT / "oracle" / "select_oracle.c")], check=True, capture_output=True)
    subprocess.run(["g++", "-O2", "-o", str(exe), str(ROOT / "tests" / "sort_check.cpp"), str(tmp_path / "sel.o")], check=True, capture_output=True)
    p = subprocess.run([str(exe), "20000"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr


def test_top1_and_topn_rules():
    sel = mapper_port.Selector()
    begin = np.array([0, 3, 3, 7, 8], np.int32)
    scores = np.array([900, 1000, 1000, 500, 700, 700, 100, 0], np.float32)
    # top1SE (ScoreBuffer.cpp:228-277): first of the equally best, numBestScore, MAPQ from best and second best
    r = sel.select_single(begin, scores)
    assert list(r["best"]) == [1, -1, 4, 7] and list(r["num_top"][[0, 2]]) == [2, 2]
    assert list(r["mapq"]) == [0, 0, 0, 0]                       # equal best scores: second best = best; score 0: no quality
    # topNSE (ScoreBuffer.cpp:279-330): sorted order, min(n, topn) alignments, MAPQ from the two best of the sorted list
    s, ns, mq, nt = mapper_port.select_topn(sel, begin, scores, 2)
    assert s.tolist() == [[1, 2], [-1, -1], [4, 5], [7, -1]] and ns.tolist() == [2, 0, 2, 1]
    assert nt.tolist()[0] == 2 and mq.tolist() == [0, 0, 0, 60]
    # strata: more equally best candidates than topn -> the read keeps none
    s, ns, mq, nt = mapper_port.select_topn(mapper_port.Selector(strata=1), np.array([0, 3], np.int32), np.array([5, 5, 5], np.float32), 2)
    assert ns.tolist() == [0] and mq.tolist() == [0] and nt.tolist() == [3]


def test_format_sam_argument_checks_and_sizing():
    """ngm_b200_format_sam is host only: NULL / shape errors are reported, an undersized buffer gets the size it needs and is left untouched."""
    import ctypes as C
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC, PAIR, SamBatch, SamOpts, _CContig, _CEncRef, load_library
    lib = load_library()
    n, stride = 4, 12
    reads = np.zeros((n, stride), np.uint8)
    reads[:, :8] = np.frombuffer(b"ACGTACGT", np.uint8)
    quals = np.where(reads != 0, ord("I"), 0).astype(np.uint8)
    names = (C.c_char_p * n)(*[b"q%d" % i for i in range(n)])
    best = np.full(n, -1, np.int32)
    zero_i, zero_f = np.zeros(n, np.int32), np.zeros(n, np.float32)
    recs = np.zeros(n, ALIGN_REC)
    pairs, heap = np.zeros(1, PAIR), np.zeros(1, np.uint8)
    ctg = (_CContig * 1)()
    ctg[0].start, ctg[0].length, ctg[0].name_len, ctg[0].name = 1000, 5000, 4, b"chr1"
    enc = _CEncRef()
    enc.concat_len, enc.n_contigs, enc.contigs = 7000, 1, C.cast(ctg, type(enc.contigs))
    so = SamOpts(0.65, 0.5, 0, 1000, 1)
    sb = SamBatch(n, stride, reads.ctypes.data, quals.ctypes.data, names, pairs.ctypes.data, zero_f.ctypes.data, best.ctypes.data, zero_i.ctypes.data,
                  zero_i.ctypes.data, None, zero_f.ctypes.data, recs.ctypes.data, heap.ctypes.data)
    used = C.c_size_t(0)
    out = np.full(4096, 0x2A, np.uint8)
    assert lib.ngm_b200_format_sam(C.byref(enc), C.byref(so), C.byref(sb), out.ctypes.data, 8, C.byref(used)) == -3      # NGM_B200_ERANGE
    need = used.value
    assert need > 8 and (out == 0x2A).all()
    assert lib.ngm_b200_format_sam(C.byref(enc), C.byref(so), C.byref(sb), out.ctypes.data, out.size, C.byref(used)) == n
    text = out[: used.value].tobytes().decode().splitlines()
    assert used.value == need and text == ["q%d\t4\t*\t0\t0\t*\t*\t0\t0\tACGTACGT\tIIIIIIII" % i for i in range(n)]      # SAMWriter.cpp:312-365
    assert lib.ngm_b200_format_sam(None, C.byref(so), C.byref(sb), out.ctypes.data, out.size, C.byref(used)) == -1
    odd = SamBatch(3, stride, reads.ctypes.data, quals.ctypes.data, names, pairs.ctypes.data, zero_f.ctypes.data, best.ctypes.data, zero_i.ctypes.data,
                   zero_i.ctypes.data, zero_i.ctypes.data, zero_f.ctypes.data, recs.ctypes.data, heap.ctypes.data)
    assert lib.ngm_b200_format_sam(C.byref(enc), C.byref(so), C.byref(odd), out.ctypes.data, out.size, C.byref(used)) == -1     # pairs need an even number of rows


def test_slam_tags_from_cigar_and_md_and_bad_strings():
    """TC:i / RA:Z / MP:Z (ngm_b200_sam_opts.slam_seq) are rebuilt from CIGAR + MD + the read (csrc/slam_tags.h): a hand-checked record, and
    records whose strings do not describe one alignment -- they come out without the tags instead of reading out of bounds."""
    import ctypes as C
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC, PAIR, SamBatch, SamOpts, _CContig, _CEncRef, load_library
    lib = load_library()
    stride = 24
    read = b"ACGTTCGATACCGT"                                     # 14 bases
    #        2S 4M 1I 3M 2D 4M : aligned columns read[2..6), read[7..10), read[10..14); MD 2A1 | 3 | ^GT | 1C2
    cases = [(b"2S4M1I3M2D4M", b"2A4^GT1C2", True), (b"2S4M1I3M2D4M", b"2A4", False), (b"2S40M", b"40", False), (b"4M", b"2^AC2", False),
             (b"14M", b"xx", False), (b"M", b"0", False), (b"14M", b"14", True)]
    n = len(cases)
    reads = np.zeros((n, stride), np.uint8)
    reads[:, : len(read)] = np.frombuffer(read, np.uint8)
    quals = np.where(reads != 0, ord("I"), 0).astype(np.uint8)
    names = (C.c_char_p * n)(*[b"q%d" % i for i in range(n)])
    pairs = np.zeros(n, PAIR)
    pairs["window_start"] = 1100
    recs = np.zeros(n, ALIGN_REC)
    heap = bytearray()
    for i, (cig, md, _) in enumerate(cases):
        recs[i]["qstart"] = 2 if cig.startswith(b"2S") else 0
        recs[i]["identity"], recs[i]["score"] = 0.9, 14.0
        recs[i]["str_off"], recs[i]["cigar_len"], recs[i]["md_len"] = len(heap), len(cig), len(md)
        heap += cig + md
    heap = np.frombuffer(bytes(heap), np.uint8).copy()
    best = np.arange(n, dtype=np.int32)
    ones, scores, mq = np.ones(n, np.int32), np.full(n, 100, np.float32), np.full(n, 60, np.int32)
    ctg = (_CContig * 1)()
    ctg[0].start, ctg[0].length, ctg[0].name_len, ctg[0].name = 1000, 5000, 4, b"chr1"
    enc = _CEncRef()
    enc.concat_len, enc.n_contigs, enc.contigs = 7000, 1, C.cast(ctg, type(enc.contigs))
    so = SamOpts(0.0, 0.0, 0, 1000, 1, 0, 0, None, 0, 1)          # slam_seq = 1
    sb = SamBatch(n, stride, reads.ctypes.data, quals.ctypes.data, names, pairs.ctypes.data, scores.ctypes.data, best.ctypes.data, mq.ctypes.data,
                  ones.ctypes.data, None, scores.ctypes.data, recs.ctypes.data, heap.ctypes.data)
    out, used = np.zeros(1 << 16, np.uint8), C.c_size_t(0)
    assert lib.ngm_b200_format_sam(C.byref(enc), C.byref(so), C.byref(sb), out.ctypes.data, out.size, C.byref(used)) == n
    lines = out[: used.value].tobytes().decode().splitlines()
    assert len(lines) == n
    for ln, (cig, md, ok) in zip(lines, cases):
        assert ("\tTC:i:" in ln) == ok, ln
        assert ln.split("\t")[5] == cig.decode() and ("MD:Z:" + md.decode()) in ln
    # the hand-checked one: 2S, then read[2..6) = G T T C against reference G T A C (MD "2A1"): one mismatch, reference A under read T
    # (type 5 * trans[A] + trans[T] = 3) at read position 5 and reference position 3 (1-based); MD's "4" counts the last column of this
    # run and the three of the next; "^GT" is the deletion; "1C2" puts a reference C under read[11]
    tags = dict(f.split(":", 2)[::2] for f in lines[0].split("\t")[11:])
    assert tags["MP"].split(",")[0] == "3:5:3"
    assert tags["TC"] == "0" and sum(int(v) for v in tags["RA"].split(",")) == 11        # 4 + 3 + 4 aligned columns
    assert "MP" not in dict(f.split(":", 2)[::2] for f in lines[-1].split("\t")[11:])     # 14M / MD 14: no mismatch, no MP tag
