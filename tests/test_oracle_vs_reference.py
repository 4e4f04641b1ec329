"""Live differential fuzz: C restatement vs the unmodified reference backend.

Needs oracle/_ref (built where /root/reference exists); skipped elsewhere -- the
committed fixtures in tests/golden/ carry the same evidence to the GPU box.
"""
import numpy as np
import pytest

from oracle import fuzzgen, port, ref_driver as rd
from tests import util

pytestmark = pytest.mark.skipif(not rd.available(), reason="oracle/_ref not built")


def scrub(refs, qrys, qml, cor):
    crefs, cqrys = fuzzgen.make_pairs(len(refs), qml, cor, 4242, clean=True)
    for _ in range(10):
        ub = np.zeros(len(refs), bool)
        for mode in (0, 1):
            ub |= np.array([a.ascore == -1.0 and a.cigar == b"!!!" for a in port.batch_align(refs, qrys, qml, cor, mode)])
        if not ub.any():
            return refs, qrys
        refs[ub], qrys[ub] = crefs[ub], cqrys[ub]
    raise AssertionError("scrub did not converge")


@pytest.mark.parametrize("qml,cor,n", [(32, 10, 1500), (102, 20, 1500), (152, 27, 1500), (252, 80, 300)])
def test_port_equals_reference_on_fresh_fuzz(qml, cor, n):
    refs, qrys = fuzzgen.make_pairs(n, qml, cor, seed=31337 + qml)
    # scores are defined for every input, including the ones the alignment path cannot handle
    for mode in (0, 1):
        r = rd.run(refs, qrys, qml, cor, mode, align=False)
        np.testing.assert_array_equal(util.bits(port.batch_score(refs, qrys, qml, cor, mode)), util.bits(r.scores))
    refs, qrys = scrub(refs, qrys, qml, cor)
    for mode in (0, 1):
        r = rd.run(refs, qrys, qml, cor, mode)
        got = [util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md)
               for a in port.batch_align(refs, qrys, qml, cor, mode)]
        want = [util.align_tuple(a.position_offset, a.qstart, a.qend, a.nm, a.identity, a.ascore, a.cigar, a.md) for a in r.aligns]
        assert got == want
