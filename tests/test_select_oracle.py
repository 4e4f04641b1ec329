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
