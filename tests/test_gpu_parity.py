"""GPU parity: the CUDA backend (through the C ABI) against the oracle and the golden fixtures.

Bit-exact on every field the reference's IAlignment returns.  Lanes that make the reference's
backtracking kernel bail out (best_read_index <= 0, oclSwCigar.cl:78) are compared against the
port's documented failure convention (Score -1, PositionOffset = best_read_index, strings untouched).
"""
import numpy as np
import pytest

from oracle import fuzzgen, port
from tests import util

pytestmark = pytest.mark.gpu


def make_sw(qml, cor, sc_kwargs=None, **kw):
    from nextgenmap_b200.host import CudaSW
    sc_kwargs = sc_kwargs or {}
    return CudaSW(qml, cor, match_bonus=sc_kwargs.get("match", 10), mismatch_penalty=sc_kwargs.get("mismatch", 15),
                  gap_read_penalty=sc_kwargs.get("gap_read", 20), gap_ref_penalty=sc_kwargs.get("gap_ref", 20),
                  match_bonus_tt=sc_kwargs.get("match_tt", 0), match_bonus_tc=sc_kwargs.get("match_tc", 0),
                  bs_mapping=sc_kwargs.get("bs_mapping", 0), slam_seq=sc_kwargs.get("slam_seq", 0), **kw)


def tuples_from_gpu(aligns):
    return [util.align_tuple(a.PositionOffset, a.QStart, a.QEnd, a.NM, a.Identity, a.Score, a.pBuffer1, a.pBuffer2) for a in aligns]


def tuples_from_port(aligns):
    return [util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md) for a in aligns]


@pytest.mark.parametrize("lane_mode", [1, 0])
@pytest.mark.parametrize("name", util.golden_names())
def test_golden_fixtures(name, lane_mode):
    g = util.load_golden(name)
    qml, cor = int(g["qml"]), int(g["corridor"])
    sw = make_sw(qml, cor, util.scoring_kwargs(g), lane_mode=lane_mode)
    dirs = g.get("dirs")
    for mode in (0, 1):
        s = sw.BatchScore(mode, g["refs"], g["qrys"], dirs)
        np.testing.assert_array_equal(util.bits(s), util.bits(g[f"score{mode}"]), err_msg=f"{name} mode {mode} scores")
        got = tuples_from_gpu(sw.BatchAlign(mode, g["refs"], g["qrys"], None, dirs))
        want = util.golden_align_tuples(g, mode)
        bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
        assert not bad, f"{name} mode {mode}: {len(bad)} differ, first {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"
    sw.close()


SHAPES = [(32, 10, 3000), (76, 16, 3000), (102, 20, 4000), (152, 27, 6000), (252, 42, 1500), (252, 80, 1000), (402, 65, 600),
          (20, 5, 500), (60, 12, 1000), (152, 25, 1000), (152, 28, 1000), (1000, 155, 40), (127, 23, 1500), (202, 35, 1000)]       # (23, 27, 35, 42: the exact-corridor forward kernels)


@pytest.mark.parametrize("lane_mode", [1, 0])
@pytest.mark.parametrize("qml,cor,n", SHAPES)
def test_fresh_fuzz_against_oracle(qml, cor, n, lane_mode):
    refs, qrys = fuzzgen.make_pairs(n, qml, cor, seed=9000 + qml * 7 + cor)
    sw = make_sw(qml, cor, lane_mode=lane_mode)
    for mode in (0, 1):
        np.testing.assert_array_equal(util.bits(sw.BatchScore(mode, refs, qrys)), util.bits(port.batch_score(refs, qrys, qml, cor, mode)),
                                      err_msg=f"scores qml {qml} corridor {cor} mode {mode}")
        got = tuples_from_gpu(sw.BatchAlign(mode, refs, qrys))
        want = tuples_from_port(port.batch_align(refs, qrys, qml, cor, mode))
        bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
        assert not bad, f"qml {qml} cor {cor} mode {mode}: {len(bad)} differ, first {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"
    sw.close()


@pytest.mark.parametrize("kw", [dict(match=5, mismatch=4, gap_read=7, gap_ref=11), dict(match=1, mismatch=3, gap_read=5, gap_ref=2),
                                dict(match=4, mismatch=2, gap_read=10, gap_ref=10, bs_mapping=1, match_tt=4, match_tc=4),
                                dict(match=10, mismatch=15, gap_read=20, gap_ref=20, slam_seq=2, match_tt=10, match_tc=15),
                                dict(match=10, mismatch=15, gap_read=20, gap_ref=20, slam_seq=1)])
def test_other_scoring_against_oracle(kw):
    qml, cor, n = 102, 20, 3000
    refs, qrys = fuzzgen.make_pairs(n, qml, cor, seed=4711)
    dirs = np.random.default_rng(3).integers(0, 2, n).astype(np.uint8) if (kw.get("bs_mapping") or kw.get("slam_seq")) else None
    sw = make_sw(qml, cor, kw)
    psc = port.Scoring(**kw)
    for mode in (0, 1):
        np.testing.assert_array_equal(util.bits(sw.BatchScore(mode, refs, qrys, dirs)), util.bits(port.batch_score(refs, qrys, qml, cor, mode, psc, dirs)))
        got = tuples_from_gpu(sw.BatchAlign(mode, refs, qrys, None, dirs))
        want = tuples_from_port(port.batch_align(refs, qrys, qml, cor, mode, psc, dirs))
        bad = [i for i, (x, y) in enumerate(zip(got, want)) if x != y]
        assert not bad, f"{kw} mode {mode}: {len(bad)} differ, first {bad[0]}: got {got[bad[0]]} want {want[bad[0]]}"
    sw.close()


def test_clip_styles():
    qml, cor, n = 102, 20, 800
    refs, qrys = fuzzgen.make_pairs(n, qml, cor, seed=12)
    for hard, silent in ((1, 0), (0, 1)):
        sw = make_sw(qml, cor, hard_clip=hard, silent_clip=silent)
        got = tuples_from_gpu(sw.BatchAlign(0, refs, qrys))
        want = tuples_from_port(port.batch_align(refs, qrys, qml, cor, 0, port.Scoring(hard_clip=hard, silent_clip=silent)))
        assert got == want
        sw.close()


def test_edge_batches():
    sw = make_sw(32, 10)
    assert len(sw.BatchScore(0, np.zeros((0, 44), np.uint8), np.zeros((0, 32), np.uint8))) == 0
    assert sw.BatchAlign(0, np.zeros((0, 44), np.uint8), np.zeros((0, 32), np.uint8)) == []
    refs, qrys = fuzzgen.make_pairs(5, 32, 10, seed=1)         # ragged quad
    np.testing.assert_array_equal(sw.BatchScore(0, refs, qrys), port.batch_score(refs, qrys, 32, 10, 0))
    from nextgenmap_b200.host import NgmB200Error
    with pytest.raises(NgmB200Error):
        sw.BatchScore(7, refs, qrys)                            # unsupported mode is an error, not a silent 0
    sw.close()
    with pytest.raises(NgmB200Error):
        make_sw(32, 10, dict(match=10.5))                       # non-integer scoring refused (SURVEY 8a note 8)


def test_large_batch_spans_chunks():
    qml, cor, n = 152, 27, 140000                                # > strict chunk (131072)
    refs, qrys = fuzzgen.make_pairs(2000, qml, cor, seed=77)
    idx = np.random.default_rng(0).integers(0, 2000, n)
    refs, qrys = refs[idx], qrys[idx]
    sw = make_sw(qml, cor)
    s = sw.BatchScore(0, refs, qrys)
    np.testing.assert_array_equal(util.bits(s), util.bits(port.batch_score(refs, qrys, qml, cor, 0)))
    sw.close()
