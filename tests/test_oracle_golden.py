"""The C restatement (oracle/ngm_oracle.c) against outputs of the reference itself.

tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the
unmodified reference backend (oracle/_ref/ngm_ref_harness).  Bit-exact on every
field: score, PositionOffset, QStart, QEnd, NM, Identity, Align.Score, CIGAR, MD.
"""
import numpy as np
import pytest

from oracle import port
from tests import util


@pytest.mark.parametrize("name", util.golden_names())
@pytest.mark.parametrize("mode", [0, 1])
def test_port_matches_reference_outputs(name, mode):
    g = util.load_golden(name)
    qml, cor = int(g["qml"]), int(g["corridor"])
    sc = port.Scoring(**util.scoring_kwargs(g))
    dirs = g.get("dirs")
    scores = port.batch_score(g["refs"], g["qrys"], qml, cor, mode, sc, dirs)
    np.testing.assert_array_equal(util.bits(scores), util.bits(g[f"score{mode}"]))
    got = [util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md)
           for a in port.batch_align(g["refs"], g["qrys"], qml, cor, mode, sc, dirs)]
    want = util.golden_align_tuples(g, mode)
    bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
    assert not bad, f"{len(bad)} alignments differ, first: {bad[0]} got {got[bad[0]]} want {want[bad[0]]}"


def test_appendix_a_known_answers():
    """SURVEY.md Appendix A rows, spelled out (local | end-free)."""
    g = util.load_golden("appendix_a")
    assert [int(s) for s in g["score0"]] == [300, 250, 280, 270, 290, 275, 50, 220, 275, 250, 200, 20, 260]
    assert [int(s) for s in g["score1"]] == [300, 250, 280, 270, 290, 275, -90, 135, 275, 175, 200, -225, 260]
    assert g["cigar0"][6] == b"5S5M20S" and g["md1"][9] == b"25xxxxx0" and g["cigar1"][6] == b"2M1I1M1I5M1I1M1I2M4I11M"
    assert g["md0"][8] == b"8N21" and int(g["nm0"][8]) == 1        # N vs N is a mismatch on the CPU device
    a = port.batch_align(g["refs"], g["qrys"], 32, 10, 0)
    assert (a[2].cigar, a[2].md, a[2].position_offset) == (b"11M1D19M", b"11^T19", 5)


def test_empty_batch_and_bad_mode():
    z = np.zeros((0, 44), np.uint8)
    assert len(port.batch_score(z, np.zeros((0, 32), np.uint8), 32, 10, 0)) == 0
    refs, qrys = np.full((4, 44), ord("A"), np.uint8), np.full((4, 32), ord("A"), np.uint8)
    out = port.batch_score(refs, qrys, 32, 10, 7)                     # unsupported mode -> nothing computed
    assert np.isnan(out).all()


def test_quad_granular_empty_read():
    """oclSwScore.cl:124 -- only lane 0 of each quad is tested for an empty read."""
    refs = np.full((8, 44), ord("A"), np.uint8)
    qrys = np.zeros((8, 32), np.uint8)
    qrys[1:4, :10] = ord("A")           # quad 0: leader empty -> whole quad inactive
    qrys[4, :10] = ord("A")             # quad 1: leader non-empty, lanes 5..7 empty -> evaluated, score 0
    s = port.batch_score(refs, qrys, 32, 10, 0)
    assert list(s) == [-1, -1, -1, -1, 100, 0, 0, 0]
    s = port.batch_score(refs, qrys, 32, 10, 1)
    assert list(s[:4]) == [-16000] * 4 and s[4] == 100 and list(s[5:]) == [0, 0, 0]


def test_decode_window_matches_sequenceprovider_rules():
    """SequenceProvider.cpp:382-441: odd offsets take the low nibble first, odd lengths end in 'x'."""
    ref = b"ACGTTGCANACGTACGTTTGACCA"
    packed = port.pack_ref(ref)
    n = len(ref)
    assert port.decode_window(packed, n, 0, 10) == b"ACGTTGCA\0\0"
    assert port.decode_window(packed, n, 1, 10) == b"CGTTGCANA\0"          # odd offset decodes len+1 bases
    assert port.decode_window(packed, n, 2, 9) == b"GTTGCANx\0"            # odd len 7: 8 chars decoded, last -> 'x'
    assert port.decode_window(packed, n, 20, 10) == b"ACCAxxxx\0\0"         # past the concatenated end -> 'x'
    assert port.decode_window(packed, n, n, 10) is None                      # offset >= length -> failure
