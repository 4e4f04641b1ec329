"""Candidate search on the device against the oracle pinned to the reference (oracle/cs_oracle.c) and
against the fixtures the reference's own code produced (tests/golden/cs): prefix table (index, weights, position table,
max_kfreq) and per-read candidate lists -- set, votes, ORDER -- through both the block-per-read kernel and the
sequential exact kernel."""
from pathlib import Path

import numpy as np
import pytest

from oracle import cs_port, port
from tests import cs_cases

pytestmark = pytest.mark.gpu

GOLD = Path(__file__).resolve().parent / "golden" / "cs"
NAMES = sorted(p.stem for p in GOLD.glob("*.npz"))


def load(name):
    with np.load(GOLD / f"{name}.npz") as z:
        return {k: z[k] for k in z.files}


def lists_from_device(begin, pairs, votes, mh, corridor):
    out = {}
    for r in range(len(begin) - 1):
        b, e = int(begin[r]), int(begin[r + 1])
        out[r] = (float(mh[r]), [((int(pairs["window_start"][j]) + (corridor >> 1)) & (2 ** 64 - 1), int(pairs["flags"][j]) & 1, float(votes[j]))
                                 for j in range(b, e)])
    return out


def shapes_for(read_len):
    return (read_len | 1) + 1, int(5 + 0.15 * read_len)


@pytest.mark.parametrize("name", NAMES)
def test_index_and_candidates_match_reference_fixture(name):
    from nextgenmap_b200.host import CudaSW
    g = load(name)
    k, read_len = int(g["k"]), int(g["read_len"])
    concat = g["concat"].tobytes()
    ctg = [(int(a), int(b)) for a, b in g["contigs"]]
    qml, cor = shapes_for(read_len)
    sw = CudaSW(qml, cor)
    sw.set_reference(port.pack_ref(concat), len(concat) - 1)
    info = sw.cs_build_index(ctg, sw.cs_params(kmer=k, sensitivity=float(g["sensitivity"])))
    assert info["table_len"] == int(g["ht_table_len"]) and info["max_kfreq"] == int(g["max_kfreq"])
    tab, weight, table = sw.cs_export_index()
    used = np.nonzero(weight != 0)[0].astype(np.uint32)
    np.testing.assert_array_equal(used, g["ht_used_prefix"])
    np.testing.assert_array_equal(weight[used], g["ht_used_weight"])
    np.testing.assert_array_equal(tab[used + 1] - tab[used], g["ht_used_count"])
    np.testing.assert_array_equal(table, g["ht_table"])
    want = {}
    for i, r in enumerate(g["read_index"]):
        b, e = int(g["cand_begin"][i]), int(g["cand_begin"][i + 1])
        want[int(r)] = (float(g["max_hit"][i]), [(int(g["cand_loc"][j]), int(g["cand_rev"][j]), float(g["cand_votes"][j])) for j in range(b, e)])
    for exact in (False, True):
        got = lists_from_device(*sw.cs_search(g["reads"], exact_only=exact), cor)
        bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
        assert not bad, f"exact={exact}: {len(bad)} reads differ, first {bad[0]}"
        if not exact:
            assert sw.cs_exact_reads() <= len(want) // 20     # the block-per-read kernel does (nearly) all of the work
    sw.close()


@pytest.mark.parametrize("seed,k,read_len,sens,scale", [(51, 13, 150, 0.5, 4), (52, 12, 100, 0.3, 2), (53, 13, 250, 0.8, 3), (54, 11, 400, 0.5, 1),
                                                        (55, 10, 36, 0.5, 1), (56, 13, 1000, 0.5, 2)])
@pytest.mark.parametrize("config", [0, 2, 3, 4])
def test_fresh_cases_against_oracle(seed, k, read_len, sens, scale, config, monkeypatch):
    """config: the library's testing hook NGM_B200_CS_CONFIG -- the small references here would always take the smallest kernel
    configuration their read length allows; 2..4 push the same cases through the configurations that long reads / large genomes use."""
    from nextgenmap_b200.host import CudaSW
    if config:
        if read_len > 250 and config < 4 or seed in (52, 55):
            pytest.skip("covered by the default configuration")
        monkeypatch.setenv("NGM_B200_CS_CONFIG", str(config))
    contigs = cs_cases.make_reference(seed, scale)
    concat, ctg, concat_len = cs_port.layout(contigs)
    qml, cor = shapes_for(read_len)
    reads = cs_cases.make_reads(seed + 1, concat, ctg, 700, read_len, qml)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=k)
    sw = CudaSW(qml, min(cor, 155))
    sw.set_reference(port.pack_ref(concat), concat_len)
    info = sw.cs_build_index(ctg, sw.cs_params(kmer=k, sensitivity=sens))
    assert info["table_len"] == ix.table_len and info["max_kfreq"] == ix.max_kfreq
    tab, weight, table = sw.cs_export_index()
    np.testing.assert_array_equal(tab, ix.tab)
    np.testing.assert_array_equal(weight, ix.weight)
    np.testing.assert_array_equal(table, ix.table)
    begin, cands, mh = ix.search(reads, sens)
    want = {r: (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
            for r in range(reads.shape[0])}
    for exact in (False, True):
        got = lists_from_device(*sw.cs_search(reads, exact_only=exact), min(cor, 155))
        bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
        assert not bad, f"exact={exact}: {len(bad)} reads differ, first {bad[0]}"
        if not exact:
            print(f"k{k} L{read_len}: {sw.cs_exact_reads()} of {len(want)} reads took the exact kernel")
            assert sw.cs_exact_reads() <= len(want) // 20
    ix.close()
    sw.close()


def test_index_loaded_from_file_arrays_equals_built_index():
    """cs_load_index with the oracle's arrays (= what NGM's ht file holds) searches like the built index."""
    from nextgenmap_b200.host import CudaSW
    contigs = cs_cases.make_reference(77)
    concat, ctg, concat_len = cs_port.layout(contigs)
    reads = cs_cases.make_reads(78, concat, ctg, 500, 100, 102)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=11)
    a, b = CudaSW(102, 20), CudaSW(102, 20)
    a.set_reference(port.pack_ref(concat), concat_len)
    a.cs_build_index(ctg, a.cs_params(kmer=11))
    info = b.cs_load_index(ix.tab, ix.weight, ix.table, b.cs_params(kmer=11))
    assert info["max_kfreq"] == ix.max_kfreq
    ra, rb = a.cs_search(reads), b.cs_search(reads)
    for x, y in zip(ra, rb):
        np.testing.assert_array_equal(x, y)
    a.close()
    b.close()


def test_candidates_feed_the_alignment_path():
    """reads -> cs_search -> score_pairs -> top-1: the true locus wins for clean reads (pipeline plumbing)."""
    from nextgenmap_b200.host import CudaSW
    rng = np.random.default_rng(3)
    seq = cs_cases.ACGT[rng.integers(0, 4, 400_000)].tobytes()
    concat, ctg, concat_len = cs_port.layout([seq])
    L, qml, cor = 150, 152, 27
    n = 2000
    reads = np.zeros((n, qml), np.uint8)
    truth = np.zeros(n, np.int64)
    for r in range(n):
        pos = int(rng.integers(0, len(seq) - L))
        s = np.frombuffer(seq[pos: pos + L], np.uint8).copy()
        mut = rng.random(L) < 0.02
        s[mut] = cs_cases.ACGT[rng.integers(0, 4, int(mut.sum()))]
        b = s.tobytes()
        if r & 1:
            b = b.translate(cs_cases.COMP)[::-1]
        reads[r, :L] = np.frombuffer(b, np.uint8)
        truth[r] = 1000 + pos
    sw = CudaSW(qml, cor)
    sw.set_reference(port.pack_ref(concat), concat_len)
    sw.cs_build_index(ctg, sw.cs_params(kmer=13))
    begin, pairs, votes, mh = sw.cs_search(reads)
    sw.set_reads(reads)
    scores = sw.score_pairs(0, pairs)
    hit = 0
    for r in range(n):
        b, e = begin[r], begin[r + 1]
        if e > b:
            j = b + int(np.argmax(scores[b:e]))
            loc = int(pairs["window_start"][j]) + (cor >> 1)
            hit += abs(loc - truth[r]) <= 8 and (int(pairs["flags"][j]) & 1) == (r & 1)
    assert hit >= 0.99 * n, hit
    sw.close()


@pytest.mark.parametrize("case", __import__("json").loads((GOLD.parent / "cs_sensitivity.json").read_text()), ids=lambda c: f"seed{c['seed']}_l{c['read_len']}")
def test_sensitivity_estimate_matches_ngm_log_and_oracle(case):
    """ReadProvider::init's estimate (NGM run without -s): device == oracle == the value the unmodified NGM logged; the per-read best
    vote with both strands added is the same through the block-per-read kernel and the sequential exact kernel."""
    import ctypes as C
    from nextgenmap_b200.host import CudaSW
    from nextgenmap_b200.host.cuda_sw import PAIR
    L = case["read_len"]
    contigs = cs_cases.make_reference(case["seed"], case["scale"])
    concat, ctg, concat_len = cs_port.layout(contigs)
    qml, cor = shapes_for(L)
    reads = cs_cases.make_reads(case["seed"] + 1, concat, ctg, case["n_reads"], L, qml)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=13)
    want, used = ix.estimate_sensitivity(reads)
    sw = CudaSW(qml, cor)
    sw.set_reference(port.pack_ref(concat), concat_len)
    sw.cs_build_index(ctg, sw.cs_params(kmer=13))
    got = sw.cs_estimate_sensitivity(reads)
    assert np.float32(got) == np.float32(want) and "%f" % got == case["ngm_logged"]
    sample = np.ascontiguousarray(reads[999::1000])
    merged = []
    for flags in (2, 3):
        begin = np.zeros(len(sample) + 1, np.int32)
        mh = np.zeros(len(sample), np.float32)
        total = C.c_size_t(0)
        pairs = np.zeros(1 << 16, dtype=PAIR)
        rc = sw.lib.ngm_b200_cs_search(sw.ctx, sample.ctypes.data, len(sample), sample.shape[1], flags, begin.ctypes.data, pairs.ctypes.data, None, 1 << 16,
                                       C.byref(total), mh.ctypes.data)
        assert rc == len(sample)
        merged.append(mh.copy())
    np.testing.assert_array_equal(merged[0], merged[1])
    # the installed estimate is what the searches now use
    begin, cands, mh = ix.search(reads[:300], want)
    b2, pairs, votes, mh2 = sw.cs_search(reads[:300])
    np.testing.assert_array_equal(begin, b2)
    np.testing.assert_array_equal(mh, mh2)
    ix.close()
    sw.close()


def test_device_to_device_index_transfer():
    """What a rank does with a prefix table that arrived over NCCL: ngm_b200_dev_cs_export_index on the sender, ngm_b200_dev_cs_load_index on the
    receiver.  The receiver's table and its candidate lists must equal the sender's."""
    import ctypes as C
    import torch
    from nextgenmap_b200.host import CudaSW
    from nextgenmap_b200.host.cuda_sw import CsParams
    contigs = cs_cases.make_reference(21)
    concat, ctg, concat_len = cs_port.layout(contigs)
    reads = cs_cases.make_reads(22, concat, ctg, 400, 100, 102)
    a, b = CudaSW(102, 20), CudaSW(102, 20)
    packed = port.pack_ref(concat)
    a.set_reference(packed, concat_len)
    b.set_reference(packed, concat_len)
    info = a.cs_build_index(ctg, a.cs_params(kmer=12))
    dev = torch.device("cuda")
    d_tab = torch.empty(info["index_len"], dtype=torch.int32, device=dev)
    d_w = torch.empty(info["index_len"], dtype=torch.int8, device=dev)
    d_t = torch.empty(max(info["table_len"], 1), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    lib = a.lib
    lib.ngm_b200_dev_cs_export_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_cs_load_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    a._check(lib.ngm_b200_dev_cs_export_index(a.ctx, d_tab.data_ptr(), d_w.data_ptr(), d_t.data_ptr(), st))
    p = CsParams(12, 2, 2, 1, 0.5, 0.0, info["max_kfreq"], 0)
    b._check(lib.ngm_b200_dev_cs_load_index(b.ctx, C.byref(p), d_tab.data_ptr(), d_w.data_ptr(), info["index_len"], d_t.data_ptr(), info["table_len"], st))
    assert b.cs_index_info() == info
    for x, y in zip(a.cs_export_index(), b.cs_export_index()):
        np.testing.assert_array_equal(x, y)
    ra, rb = a.cs_search(reads), b.cs_search(reads)
    for x, y in zip(ra, rb):
        np.testing.assert_array_equal(x, y)
    a.close()
    b.close()


# -- bs-mapping / SLAMseq candidate search (CS::PrefixMutateSearch, CS.cpp:53-112) ---------------------------------
GOLD_MUT = GOLD.parent / "cs_mut"
NAMES_MUT = sorted(p.stem for p in GOLD_MUT.glob("*.npz"))


def _mut_context(g_or_none, concat, ctg, concat_len, k, read_len, sens, mode, paired, bs_cutoff, max_kfreq=0):
    from nextgenmap_b200.host import CudaSW
    qml, cor = shapes_for(read_len)
    sw = CudaSW(qml, cor, bs_mapping=1 if mode == 1 else 0, slam_seq=4 if mode == 2 else 0)
    sw.set_reference(port.pack_ref(concat), concat_len)
    # CompactPrefixTable indexes every reference position under bs_mapping (PrefixTable.cpp:204-207); the reads' k-mers are thinned instead
    info = sw.cs_build_index(ctg, sw.cs_params(kmer=k, kmer_skip=0 if mode == 1 else 2, sensitivity=sens, max_kfreq=max_kfreq))
    sw.cs_configure_mutation(bs_mapping=1 if mode == 1 else 0, slam_seq=4 if mode == 2 else 0, bs_cutoff=bs_cutoff, paired=paired, read_kmer_skip=2)
    return sw, cor, info


@pytest.mark.parametrize("name", NAMES_MUT)
def test_mutated_search_matches_reference_fixture(name):
    """tests/golden/cs_mut: what the reference's own CS code produced under --bs-mapping / --slam-seq 4 (single-end and paired)."""
    with np.load(GOLD_MUT / f"{name}.npz") as z:
        g = {k: z[k] for k in z.files}
    concat = g["concat"].tobytes()
    ctg = [(int(a), int(b)) for a, b in g["contigs"]]
    sw, cor, info = _mut_context(g, concat, ctg, len(concat) - 1, int(g["k"]), int(g["read_len"]), float(g["sensitivity"]), int(g["mode"]), bool(g["paired"]),
                                 int(g["bs_cutoff"]))
    assert info["max_kfreq"] == int(g["max_kfreq"])
    want = {}
    for i, r in enumerate(g["read_index"]):
        b, e = int(g["cand_begin"][i]), int(g["cand_begin"][i + 1])
        want[int(r)] = (float(g["max_hit"][i]), [(int(g["cand_loc"][j]), int(g["cand_rev"][j]), float(g["cand_votes"][j])) for j in range(b, e)])
    begin, pairs, votes, mh = sw.cs_search(g["reads"])
    got = lists_from_device(begin, pairs, votes, mh, cor)
    bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
    assert not bad, f"{len(bad)} reads differ, first {bad[0]}"
    assert sw.cs_exact_reads() == len(want)                      # the mutated k-mers are enumerated by the sequential kernel
    # the direction flag ScoreBuffer would hand to BatchScore (ScoreBuffer.cpp:92-110): the strand, inverted for second mates
    rd = np.repeat(np.arange(len(begin) - 1), np.diff(begin))
    rev = (pairs["flags"] & 1).astype(bool)
    second = (rd & 1).astype(bool) if int(g["paired"]) else np.zeros(len(rd), bool)
    np.testing.assert_array_equal(((pairs["flags"] >> 1) & 1).astype(bool), rev ^ second)
    # switching the mutation off again gives the plain search
    sw.cs_configure_mutation()
    ix = cs_port.Index(port.pack_ref(concat), len(concat) - 1, ctg, k=int(g["k"]), ref_skip=int(g["ref_skip"]))
    b2, c2, m2 = ix.search(g["reads"], float(g["sensitivity"]), max_kfreq=int(g["max_kfreq"]))
    plain = {r: (float(m2[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in c2[b2[r]: b2[r + 1]]]) for r in range(g["reads"].shape[0])}
    assert lists_from_device(*sw.cs_search(g["reads"]), cor) == plain
    ix.close()
    sw.close()


@pytest.mark.parametrize("seed,k,read_len,sens,mode,paired,cutoff", [(61, 13, 150, 0.5, 1, False, 6), (62, 11, 100, 0.4, 1, True, 4), (63, 13, 150, 0.5, 2, False, 6),
                                                                      (64, 12, 250, 0.7, 2, True, 6), (65, 10, 60, 0.5, 1, False, 8)])
def test_mutated_search_fresh_cases_against_oracle(seed, k, read_len, sens, mode, paired, cutoff):
    contigs = cs_cases.make_reference(seed, 2)
    concat, ctg, concat_len = cs_port.layout(contigs)
    qml, cor = shapes_for(read_len)
    reads = cs_cases.convert_bases(cs_cases.make_reads(seed + 1, concat, ctg, 500, read_len, qml), seed + 2, mode, paired)
    ix = cs_port.Index(port.pack_ref(concat), concat_len, ctg, k=k, ref_skip=0 if mode == 1 else 2)
    sw, cor, info = _mut_context(None, concat, ctg, concat_len, k, read_len, sens, mode, paired, cutoff)
    assert info["table_len"] == ix.table_len and info["max_kfreq"] == ix.max_kfreq
    begin, cands, mh = ix.search_mut(reads, sens, mode, bs_cutoff=cutoff, paired=paired, read_skip=2 if mode == 1 else 0)
    want = {r: (float(mh[r]), [(int(c["location"]), int(c["reverse"]), float(c["score"])) for c in cands[begin[r]: begin[r + 1]]])
            for r in range(reads.shape[0])}
    got = lists_from_device(*sw.cs_search(reads), cor)
    bad = [(r, want[r], got[r]) for r in want if want[r] != got[r]]
    assert not bad, f"{len(bad)} reads differ, first {bad[0]}"
    assert sum(len(v[1]) for v in want.values()) > 300
    ix.close()
    sw.close()
