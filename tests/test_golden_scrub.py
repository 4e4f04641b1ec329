"""How much of every golden fixture was chosen by the port rather than by the random generator.

tests/golden/make_golden.py replaces pairs on which the reference's backtracking kernel is skipped (oclSwCigar.cl:78: its host code then
reads uninitialised memory) by clean pairs BEFORE the reference sees them; the port decides which.  A port bug that mis-flagged legal
pairs would hide them from the reference for ever, so this test replays the selection from the committed seeds, checks that the committed
inputs are exactly its result, and bounds the number of replaced pairs per fixture."""
import numpy as np
import pytest

from oracle import fuzzgen, port
from tests import util
from tests.golden import make_golden as mg

SHAPES = [(32, 10, 400), (76, 16, 300), (102, 20, 400), (152, 27, 400), (252, 42, 160), (252, 80, 120), (402, 65, 100)]
SCORINGS = [("custom_5_4_7_11", dict(match=5, mismatch=4, gap_read=7, gap_ref=11)),
            ("bs_4_2_10_10", dict(match=4, mismatch=2, gap_read=10, gap_ref=10, bs_mapping=1, match_tt=4, match_tc=4)),
            ("slam_10_15_20_20", dict(match=10, mismatch=15, gap_read=20, gap_ref=20, slam_seq=2, match_tt=10, match_tc=15))]
MAX_REPLACED = 0.20          # measured: 5 .. 16 % (the generator makes short / empty / unalignable reads on purpose)


def replaced(raw_refs, raw_qrys, g):
    same = (raw_refs == g["refs"]).all(axis=1) & (raw_qrys == g["qrys"]).all(axis=1)
    return int((~same).sum())


@pytest.mark.parametrize("qml,cor,n", SHAPES)
def test_default_scoring_fixture_inputs(qml, cor, n):
    g = util.load_golden(f"fuzz_q{qml}_c{cor}")
    raw_refs, raw_qrys = fuzzgen.make_pairs(n, qml, cor, 20261017 + qml * 100 + cor)
    refs, qrys = mg.scrub(raw_refs.copy(), raw_qrys.copy(), qml, cor, port.Scoring(), None)
    np.testing.assert_array_equal(refs, g["refs"])          # the committed inputs are the replayed selection
    np.testing.assert_array_equal(qrys, g["qrys"])
    k = replaced(raw_refs, raw_qrys, g)
    print(f"fuzz_q{qml}_c{cor}: {k} of {n} pairs replaced before the reference ran")
    assert k <= MAX_REPLACED * n


@pytest.mark.parametrize("name,kw", SCORINGS)
def test_other_scoring_fixture_inputs(name, kw):
    qml, cor, n = 102, 20, 300
    g = util.load_golden(name)
    raw_refs, raw_qrys = fuzzgen.make_pairs(n, qml, cor, 777 + len(name))
    dirs = np.random.default_rng(5).integers(0, 2, n).astype(np.uint8) if ("bs" in name or "slam" in name) else None
    refs, qrys = mg.scrub(raw_refs.copy(), raw_qrys.copy(), qml, cor, port.Scoring(**kw), dirs)
    np.testing.assert_array_equal(refs, g["refs"])
    np.testing.assert_array_equal(qrys, g["qrys"])
    k = replaced(raw_refs, raw_qrys, g)
    print(f"{name}: {k} of {n} pairs replaced before the reference ran")
    assert k <= MAX_REPLACED * n
