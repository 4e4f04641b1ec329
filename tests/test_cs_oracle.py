"""Candidate search: the C restatement (oracle/cs_oracle.c) against what the UNMODIFIED reference produced --
committed fixtures (tests/golden/cs, made by tests/golden/make_cs_golden.py) and, where oracle/_ref/ngm/ngm_cs_probe
exists, a fresh differential run."""
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import cs_port, port
from tests import cs_cases

GOLD = Path(__file__).resolve().parent / "golden" / "cs"
NAMES = sorted(p.stem for p in GOLD.glob("*.npz"))


def load(name):
    with np.load(GOLD / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def golden_lists(g):
    out = {}
    for i, r in enumerate(g["read_index"]):
        b, e = int(g["cand_begin"][i]), int(g["cand_begin"][i + 1])
        out[int(r)] = (float(g["max_hit"][i]), [(int(g["cand_loc"][j]), int(g["cand_rev"][j]), float(g["cand_votes"][j])) for j in range(b, e)])
    return out


def oracle_lists(ix, reads, sens, max_kfreq):
    begin, cands, mh = ix.search(reads, sens, max_kfreq=max_kfreq)
    return {r: (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
            for r in range(reads.shape[0])}


@pytest.mark.parametrize("name", NAMES)
def test_index_and_candidates_match_reference_fixture(name):
    g = load(name)
    k = int(g["k"])
    concat = g["concat"].tobytes()
    ctg = [(int(a), int(b)) for a, b in g["contigs"]]
    ix = cs_port.Index(port.pack_ref(concat), len(concat) - 1, ctg, k=k)
    # the prefix table the reference wrote: used prefixes, their weights, list lengths and the whole position table
    used = np.nonzero(ix.weight != 0)[0].astype(np.uint32)
    np.testing.assert_array_equal(used, g["ht_used_prefix"])
    np.testing.assert_array_equal(ix.weight[used], g["ht_used_weight"])
    np.testing.assert_array_equal(ix.tab[used + 1] - ix.tab[used], g["ht_used_count"])
    assert ix.table_len == int(g["ht_table_len"]) and ix.max_kfreq == int(g["max_kfreq"])
    np.testing.assert_array_equal(ix.table, g["ht_table"])
    want = golden_lists(g)
    got = oracle_lists(ix, g["reads"], float(g["sensitivity"]), int(g["max_kfreq"]))
    bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
    assert not bad, f"{len(bad)} reads differ, first {bad[0]}"
    assert sum(1 for r in want if len(want[r][1]) > 1) > 10       # the order of multi-candidate lists is part of the check
    ix.close()


@pytest.mark.skipif(not cs_port.probe_available(), reason="oracle/_ref/ngm/ngm_cs_probe not built")
@pytest.mark.parametrize("seed,k,read_len,sens", [(7, 10, 75, 0.5), (8, 11, 120, 0.9)])
def test_fresh_differential_run_against_reference(seed, k, read_len, sens):
    contigs = cs_cases.make_reference(seed)
    concat, ctg, concat_len = cs_port.layout(contigs)
    reads = cs_cases.make_reads(seed + 1, concat, ctg, 600, read_len, (read_len | 1) + 1)
    with tempfile.TemporaryDirectory(prefix="csdiff_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        head, rows = cs_port.run_probe(d, "ref.fa", "reads.fq", sens, k=k)
        ht = cs_port.read_ht_file(d / f"ref.fa-ht-{k}-2.3.ngm")
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=k)
    np.testing.assert_array_equal(ix.tab, ht["tab"])
    np.testing.assert_array_equal(ix.weight, ht["weight"])
    np.testing.assert_array_equal(ix.table, ht["table"])
    assert ix.max_kfreq == head["max_kfreq"]
    got = oracle_lists(ix, reads, sens, head["max_kfreq"])
    for (rid, name, ln, mh, cl) in rows:
        r = int(name[1:])
        assert got[r] == (mh, cl), (r, got[r], (mh, cl))
    ix.close()


def test_revcomp_prefix():
    # A0 C1 T2 G3: ACGTT -> revcomp AACGT
    enc = {"A": 0, "C": 1, "T": 2, "G": 3}
    code = lambda s: sum(enc[c] << (2 * (len(s) - 1 - i)) for i, c in enumerate(s))
    assert cs_port.revcomp_prefix(code("ACGTT"), 5) == code("AACGT")
    assert cs_port.revcomp_prefix(code("ACGTTGCATGCAA"), 13) == code("TTGCATGCAACGT")


SENS = Path(__file__).resolve().parent / "golden" / "cs_sensitivity.json"


@pytest.mark.parametrize("case", __import__("json").loads(SENS.read_text()), ids=lambda c: f"seed{c['seed']}_l{c['read_len']}")
def test_sensitivity_estimate_matches_ngm_log(case):
    """ReadProvider::init's estimate (no -s): the restatement against the value the unmodified NGM logged (tests/golden/make_cs_sensitivity_golden.py)."""
    contigs = cs_cases.make_reference(case["seed"], case["scale"])
    concat, ctg, concat_len = cs_port.layout(contigs)
    reads = cs_cases.make_reads(case["seed"] + 1, concat, ctg, case["n_reads"], case["read_len"], (case["read_len"] | 1) + 1)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=13)
    sens, used = ix.estimate_sensitivity(reads)
    assert "%f" % sens == case["ngm_logged"] and used > 0
    ix.close()


def test_sensitivity_estimate_needs_1000_reads():
    contigs = cs_cases.make_reference(5)
    concat, ctg, concat_len = cs_port.layout(contigs)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=13)
    assert ix.estimate_sensitivity(cs_cases.make_reads(6, concat, ctg, 999, 100, 102)) == (0.5, 0)
    ix.close()


# -- bs-mapping / SLAMseq: CS::PrefixMutateSearch (CS.cpp:53-112) ------------------------------------------------------------------
GOLD_MUT = GOLD.parent / "cs_mut"
NAMES_MUT = sorted(p.stem for p in GOLD_MUT.glob("*.npz"))


def load_mut(name):
    with np.load(GOLD_MUT / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def oracle_lists_mut(ix, g, paired=None):
    begin, cands, mh = ix.search_mut(g["reads"], float(g["sensitivity"]), int(g["mode"]), bs_cutoff=int(g["bs_cutoff"]),
                                     paired=bool(g["paired"]) if paired is None else paired, read_skip=int(g["read_skip"]), max_kfreq=int(g["max_kfreq"]))
    return {r: (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
            for r in range(g["reads"].shape[0])}


@pytest.mark.parametrize("name", NAMES_MUT)
def test_mutated_search_matches_reference_fixture(name):
    """Fixtures made by tests/golden/make_cs_mut_golden.py from the reference's own CS code under --bs-mapping / --slam-seq 4."""
    g = load_mut(name)
    concat = g["concat"].tobytes()
    ctg = [(int(a), int(b)) for a, b in g["contigs"]]
    ix = cs_port.Index(port.pack_ref(concat), len(concat) - 1, ctg, k=int(g["k"]), ref_skip=int(g["ref_skip"]))
    want = golden_lists(g)
    got = oracle_lists_mut(ix, g)
    bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
    assert not bad, f"{len(bad)} reads differ, first {bad[0]}"
    # the fixture must exercise what it claims: the mutation changes the lists, and the mates mutate different bases
    plain = oracle_lists(ix, g["reads"], float(g["sensitivity"]), int(g["max_kfreq"]))
    assert sum(1 for r in want if want[r] != plain[r]) > 20
    if int(g["paired"]):
        single = oracle_lists_mut(ix, g, paired=False)
        assert sum(1 for r in want if (r & 1) and want[r] != single[r]) > 10
    if int(g["mode"]) == 2:
        assert sum(1 for r in want for c in want[r][1] if c[2] != int(c[2])) > 50          # fractional votes (weight 1 / m_CurrentMutLocs)
    ix.close()


@pytest.mark.skipif(not cs_port.probe_available(), reason="oracle/_ref/ngm/ngm_cs_probe not built")
@pytest.mark.parametrize("seed,k,read_len,sens,mode,paired,opts", [(17, 11, 90, 0.5, 1, False, {}), (18, 13, 120, 0.4, 2, True, {}),
                                                                   (19, 12, 100, 0.5, 1, True, {"kmer_skip": 1, "bs_cutoff": 2}),
                                                                   (20, 10, 60, 0.7, 1, False, {"kmer_skip": 4, "bs_cutoff": 9})])
def test_fresh_mutated_differential_run_against_reference(seed, k, read_len, sens, mode, paired, opts):
    """opts: under --bs-mapping `--kmer-skip` thins the READ's k-mers (CS.cpp:556-560; the index keeps every position) and `--bs-cutoff`
    bounds the replaceable bases per k-mer."""
    contigs = cs_cases.make_reference(seed)
    concat, ctg, concat_len = cs_port.layout(contigs)
    reads = cs_cases.convert_bases(cs_cases.make_reads(seed + 1, concat, ctg, 300, read_len, (read_len | 1) + 1), seed + 2, mode, paired)
    extra = (["--bs-mapping"] if mode == 1 else ["--slam-seq", "4"]) + (["-p", "--skip-mate-check"] if paired else [])
    read_skip, cutoff = opts.get("kmer_skip", 2), opts.get("bs_cutoff", 6)
    if opts:
        extra += ["--kmer-skip", str(read_skip), "--bs-cutoff", str(cutoff)]
    with tempfile.TemporaryDirectory(prefix="csmutdiff_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        head, rows = cs_port.run_probe(d, "ref.fa", "reads.fq", sens, k=k, extra=extra)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=k, ref_skip=0 if mode == 1 else 2)
    begin, cands, mh = ix.search_mut(reads, sens, mode, bs_cutoff=cutoff, paired=paired, read_skip=read_skip if mode == 1 else 0, max_kfreq=head["max_kfreq"])
    f32 = lambda x: float(np.float32(x))
    for (rid, name, ln, m, cl) in rows:
        r = int(name[1:])
        got = (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
        assert got == (f32(m), [(a, b, f32(v)) for a, b, v in cl]), (r, got, (m, cl))
    ix.close()


@pytest.mark.skipif(not cs_port.probe_available(), reason="oracle/_ref/ngm/ngm_cs_probe not built")
def test_mutated_search_table_overflow_retries_match_reference():
    """CS::RunBatch repeats a read's search in a larger table when the 2^16-slot table runs out of probe steps (CS.cpp:386-430): bs-mapping on
    a 30 Mbp reference with k 10 gives a 250 bp read ~35 000 distinct bins (one retry, 2^18 slots) and a 900 bp read ~410 000 (three retries,
    up to 2^20 slots).  The restatement must take the same decisions -- a search that is cut short or repeated once too often changes votes."""
    import ctypes as C
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    ref_len, k = 30_000_000, 10
    contigs = [acgt[rng.integers(0, 4, ref_len)].tobytes()]
    concat, ctg, concat_len = cs_port.layout(contigs)
    full = np.frombuffer(concat, np.uint8)
    lens = [250] * 16 + [900] * 3
    reads = np.zeros((len(lens), 902), np.uint8)
    for r, ln in enumerate(lens):
        pos = int(rng.integers(2000, ref_len - 2000))
        s = full[pos: pos + ln].copy()
        s[(s == ord("C")) & (rng.random(ln) < 0.9)] = ord("T")
        reads[r, :ln] = s
    with tempfile.TemporaryDirectory(prefix="csovf_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        head, rows = cs_port.run_probe(d, "ref.fa", "reads.fq", 0.5, k=k, extra=["--bs-mapping"])
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=k, ref_skip=0)
    begin, cands, mh = ix.search_mut(reads, 0.5, 1, read_skip=2, max_kfreq=head["max_kfreq"])
    lib = port.lib()
    lib.cs_oracle_last_retries.restype = C.c_longlong
    lib.cs_oracle_last_dropped.restype = C.c_longlong
    assert lib.cs_oracle_last_retries() >= 20 and lib.cs_oracle_last_dropped() == 0      # (24 here: nearly every read is searched again)
    for (rid, name, ln, m, cl) in rows:
        r = int(name[1:])
        got = (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
        assert got == (m, cl), (r, got, (m, cl))
        assert len(cl) >= 1 and cl[0][2] > 30
    ix.close()
