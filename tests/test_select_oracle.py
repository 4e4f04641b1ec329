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
