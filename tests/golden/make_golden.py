"""Generate tests/golden/*.npz by running the UNMODIFIED reference backend.

Run in a container where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

Each fixture holds seeded inputs (windows, reads) and what the reference's
``SWOclCigar::BatchScore`` / ``BatchAlign`` (CPU OpenCL device) returned for
them in local (mode 0) and end-free (mode 1) mode.  Pairs whose alignment
would make the reference read uninitialised memory (backtracking skipped,
oclSwCigar.cl:78 -- the reference crashes or returns garbage there) are
replaced by well-defined pairs before the reference is run; the scrub uses
the port only to *select inputs*, every stored output comes from the reference.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import fuzzgen, port, ref_driver as rd  # noqa: E402

OUT = Path(__file__).resolve().parent

W = b"GCCCAGTGTGAATCGCTTAAGGGTTAAGTAAGTGTGATGCAT"
R = b"GTGTGAATCGCTTAAGGGTTAAGTAAGTGT"
WN = b"GCCCAGTGTGAATNGCTTAAGGGTTAAGTAAGTGTGATGCAT"
APPENDIX_A = [  # SURVEY.md Appendix A, qml 32 / corridor 10
    (W, R), (W, b"GTGTGCATCGCTTAAGGGTTGAGTAAGTGT"), (W, b"GTGTGAATCGCTAAGGGTTAAGTAAGTGTG"),
    (W, b"GTGTGAATCGCTTAATGGGTTAAGTAAGTG"), (W, b"GTGTGAATCGNTTAAGGGTTAAGTAAGTGT"), (WN, R),
    (W, b"GACACTCGCTATGAATCTCTGATTTACCCA"), (W, b"CTCTGAATCGCTTAAGGGTTAAGTAGCCAA"),
    (WN, b"GTGTGAATNGCTTAAGGGTTAAGTAAGTGT"), (b"GCCCAGTGTGAATCGCTTAAGGGTTAAGTAxxxxxxxxxxxx", R),
    (W, b"GTGTGAATCGCTTAAGGGTT"), (W, b"A" * 30), (W, b"GTGTGAATCGTAAGGGTTAAGTAAGTGTGA"),
]


def scrub(refs, qrys, qml, cor, psc, dirs):
    crefs, cqrys = fuzzgen.make_pairs(len(refs), qml, cor, 4242, clean=True)
    for _ in range(10):
        ub = np.zeros(len(refs), bool)
        for mode in (0, 1):
            a = port.batch_align(refs, qrys, qml, cor, mode, psc, dirs)
            ub |= np.array([x.ascore == -1.0 and x.cigar == b"!!!" for x in a])
        if not ub.any():
            return refs, qrys
        refs[ub] = crefs[ub]
        qrys[ub] = cqrys[ub]
    raise RuntimeError("scrub did not converge")


def strs(items):
    return np.array(items, dtype=object)


def emit(name, refs, qrys, qml, cor, sc_kwargs=None, dirs=None):
    sc_kwargs = sc_kwargs or {}
    rsc = rd.Scoring(**sc_kwargs)
    data = dict(refs=refs, qrys=qrys, qml=qml, corridor=cor,
                scoring=np.array([rsc.match, rsc.mismatch, rsc.gap_read, rsc.gap_ref, rsc.bs_mapping, rsc.slam_seq,
                                  rsc.match_tt, rsc.match_tc], dtype=np.int32))
    if dirs is not None:
        data["dirs"] = dirs
    for mode in (0, 1):
        r = rd.run(refs, qrys, qml, cor, mode, sc=rsc, dirs=dirs)
        data[f"score{mode}"] = r.scores
        data[f"pos{mode}"] = np.array([a.position_offset for a in r.aligns], np.int32)
        data[f"qstart{mode}"] = np.array([a.qstart for a in r.aligns], np.int32)
        data[f"qend{mode}"] = np.array([a.qend for a in r.aligns], np.int32)
        data[f"nm{mode}"] = np.array([a.nm for a in r.aligns], np.int32)
        data[f"identity{mode}"] = np.array([a.identity for a in r.aligns], np.float32)
        data[f"ascore{mode}"] = np.array([a.ascore for a in r.aligns], np.float32)
        data[f"cigar{mode}"] = np.array([a.cigar for a in r.aligns], dtype="S")
        data[f"md{mode}"] = np.array([a.md for a in r.aligns], dtype="S")
    np.savez_compressed(OUT / f"{name}.npz", **data)
    print(f"wrote {name}.npz: {len(refs)} pairs, qml {qml}, corridor {cor}")


def main():
    if not rd.available():
        raise SystemExit("oracle/_ref not built (make -C oracle ref)")
    qml, cor = 32, 10
    rbl = fuzzgen.ref_buf_len_score(qml, cor)
    emit("appendix_a", rd.pack_fixed([c[0] for c in APPENDIX_A], rbl), rd.pack_fixed([c[1] for c in APPENDIX_A], qml), qml, cor)
    shapes = [(32, 10, 400), (76, 16, 300), (102, 20, 400), (152, 27, 400), (252, 42, 160), (252, 80, 120), (402, 65, 100)]
    for qml, cor, n in shapes:
        refs, qrys = fuzzgen.make_pairs(n, qml, cor, 20261017 + qml * 100 + cor)
        refs, qrys = scrub(refs, qrys, qml, cor, port.Scoring(), None)
        emit(f"fuzz_q{qml}_c{cor}", refs, qrys, qml, cor)
    # non-default and ALT scoring (bs-mapping / SLAMseq), SURVEY 8a note 11
    for name, kw in [("custom_5_4_7_11", dict(match=5, mismatch=4, gap_read=7, gap_ref=11)),
                     ("bs_4_2_10_10", dict(match=4, mismatch=2, gap_read=10, gap_ref=10, bs_mapping=1, match_tt=4, match_tc=4)),
                     ("slam_10_15_20_20", dict(match=10, mismatch=15, gap_read=20, gap_ref=20, slam_seq=2, match_tt=10, match_tc=15))]:
        qml, cor, n = 102, 20, 300
        refs, qrys = fuzzgen.make_pairs(n, qml, cor, 777 + len(name))
        dirs = None
        if "bs" in name or "slam" in name:
            dirs = np.random.default_rng(5).integers(0, 2, n).astype(np.uint8)
        pkw = {k: v for k, v in kw.items()}
        refs, qrys = scrub(refs, qrys, qml, cor, port.Scoring(**pkw), dirs)
        emit(name, refs, qrys, qml, cor, kw, dirs)


if __name__ == "__main__":
    main()
