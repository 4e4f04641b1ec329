"""Whole mapping runs on the CPU restatements (oracle/mapper_port.py: candidate search -> scores -> selection -> alignments, rendered by
the host mirror of SAMWriter) against the UNMODIFIED NextGenMap: single-end and paired-end (`-p -t 1`; BASELINE configs[2] shape).
This is what pins oracle/select_oracle.c -- ScoreBuffer::top1SE / top1PE / CheckPairs incl. the order std::sort leaves equal scores in
and the running insert-size mean that breaks ties between equally scoring pairs (the paired inputs hold repeated segments for that).

The golden SAM (tests/golden/pe_l100.sam.gz, made by tests/golden/make_pe_golden.py from the reference binary) is checked always; the
live runs need oracle/_ref/ngm/ngm_ref."""
import gzip
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import cs_port, mapper_port, ngm_e2e as e2e, port

GOLD = Path(__file__).resolve().parent / "golden"


def read_fastq(path, paired):
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        nm = lines[i][1:].decode().split()[0]
        if paired and len(nm) > 2 and nm[-2] == "/":        # ReadProvider::NextRead strips the mate suffix (ReadProvider.cpp:417-420)
            nm = nm[:-2]
        names.append(nm)
        seqs.append(lines[i + 1])
        quals.append(lines[i + 3])
    return names, seqs, quals


def rows(seqs, qml):
    reads = np.zeros((len(seqs), qml), np.uint8)
    for i, s in enumerate(seqs):
        reads[i, : len(s)] = np.frombuffer(s, np.uint8)
    return reads


class LaidOutReference:
    """The concatenated reference as SequenceProvider::Init lays it out (SequenceProvider.cpp:289-330) from a FASTA file, with convert()
    (SequenceProvider.cpp:111-141) -- stands in for NGM's `<ref>-enc.2.ngm` where the reference binary is not available."""

    def __init__(self, fasta: Path):
        from oracle import port
        names, seqs = [], []
        for ln in open(fasta, "rb").read().split(b"\n"):
            if ln.startswith(b">"):
                names.append(ln[1:].split()[0].decode())
                seqs.append([])
            elif ln:
                seqs[-1].append(ln)
        concat, ctg, self.concat_len = cs_port.layout([b"".join(s) for s in seqs])
        self.packed = port.pack_ref(concat)
        self.contigs = [(nm, st, ln) for nm, (st, ln) in zip(names, ctg)]
        self.starts = [st for _, st, _ in self.contigs] + [len(concat) + 1000]

    def convert(self, pos: int):
        import bisect
        i = bisect.bisect_right(self.starts, pos)            # std::upper_bound over refStartPos
        if i >= len(self.starts) or self.starts[i] - pos < 1000 or i == 0:
            return None
        return i - 1, pos - self.starts[i - 1]

    def as_encoded_reference(self):
        """The same reference as the C struct ngm_b200_format_sam takes (ngm_b200_encref)."""
        import ctypes as C
        from types import SimpleNamespace
        from nextgenmap_b200.host.cuda_sw import _CContig, _CEncRef
        arr = (_CContig * len(self.contigs))()
        for i, (nm, st, ln) in enumerate(self.contigs):
            arr[i].start, arr[i].length, arr[i].name_len, arr[i].name = st, ln, len(nm), nm.encode()
        c = _CEncRef()
        c.concat_len, c.packed_bytes, c.n_contigs = self.concat_len, len(self.packed), len(self.contigs)
        c.packed = self.packed.ctypes.data_as(type(c.packed))
        c.contigs = C.cast(arr, type(c.contigs))
        return SimpleNamespace(c=c, _keep=(arr, self.packed))

    def close(self):
        pass


def oracle_sam(d: Path, read_len: int, mode: int, sens: float, paired: bool, batches: int = 1, sel=None, limits=None, mutate=None, scoring=None):
    from nextgenmap_b200.host import EncodedReference        # host-only reader of <ref>-enc.2.ngm (no CUDA call)
    from nextgenmap_b200.host import pipeline
    enc = d / "ref.fa-enc.2.ngm"
    ref = EncodedReference(str(enc)) if enc.exists() else LaidOutReference(d / "ref.fa")
    names, seqs, quals = read_fastq(d / "reads.fq", paired)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = rows(seqs, qml)
    ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13, ref_skip=0 if (mutate and mutate["mode"] == 1) else 2)
    sel = sel or mapper_port.Selector()
    got, n = [], len(names)
    step = (n // batches + 1) & ~1
    for lo in range(0, n, step):                             # the insert-size sums carry over the batches
        hi = min(n, lo + step)
        batch = mapper_port.map_batch(ref.packed, ref.concat_len, ix, reads[lo:hi], qml, cor, mode, sens, sel, paired=paired, mutate=mutate, scoring=scoring)
        if scoring is not None and scoring.bs_mapping == 1:
            batch.bs_mapping = 1                             # ZS:Z in the SAM record
        if scoring is not None and scoring.slam_seq != 0:
            batch.slam_seq = scoring.slam_seq                # TC:i / RA:Z / MP:Z
        if scoring is not None and batch.best_pair.size:     # the native formatter gives the same lines (host only)
            fmt = pipeline.format_sam(with_heap(batch), reads[lo:hi], names[lo:hi], quals[lo:hi],
                                      ref if enc.exists() else ref.as_encoded_reference(), paired, bs_mapping=getattr(batch, "bs_mapping", 0),
                                      slam_seq=getattr(batch, "slam_seq", 0), **(limits or {})).decode().splitlines()
            assert fmt == (pipeline.sam_lines_paired(batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, cor, **(limits or {})) if paired
                           else pipeline.sam_lines(None, batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, cor))
        if paired:
            got += pipeline.sam_lines_paired(batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, cor, **(limits or {}))
        else:
            got += pipeline.sam_lines(None, batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, cor)
    ix.close()
    ref.close()
    return sorted(got), sel


def diff(got, want):
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"


def test_paired_golden_sam():
    """Inputs regenerated from the seed; NGM's answer from the committed fixture."""
    with tempfile.TemporaryDirectory(prefix="pegold_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=300_000, n_frags=600, read_len=100, seed=77)
        got, sel = oracle_sam(d, 100, 0, 0.5, True, batches=3)
    want = gzip.open(GOLD / "pe_l100.sam.gz", "rt").read().splitlines()
    diff(got, want)
    flags = {ln.split("\t")[1] for ln in want}
    assert {"99", "147", "83", "163"} <= flags and len(flags) >= 12          # proper pairs, broken pairs, unmapped mates
    assert sel.state.dist_count > 300


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("ref_len,n_frags,read_len,seed,extra", [(600_000, 1500, 100, 5, []), (500_000, 1000, 150, 6, ["-e"]), (400_000, 800, 75, 7, ["--fast-pairing"]),
                                                                 (400_000, 900, 100, 61, ["-I", "250", "-X", "430"]), (400_000, 900, 100, 62, ["--strata"])])
def test_paired_sam_identical_to_ngm(ref_len, n_frags, read_len, seed, extra):
    with tempfile.TemporaryDirectory(prefix="pe_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=ref_len, n_frags=n_frags, read_len=read_len, seed=seed)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["-p", "-s", "0.5", *extra]) if not ln.startswith("@")]
        limits = {"min_insert_size": int(extra[extra.index("-I") + 1]), "max_insert_size": int(extra[extra.index("-X") + 1])} if "-I" in extra else {}
        sel = mapper_port.Selector(fast_pairing=1 if "--fast-pairing" in extra else 0, strata=1 if "--strata" in extra else 0, **limits)
        got, _ = oracle_sam(d, read_len, 1 if "-e" in extra else 0, 0.5, True, batches=2, sel=sel, limits=limits)
    diff(got, want)


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("extra,filters,sub_rate,indels", [([], {}, 0.02, 0.15), (["-i", "0.93", "-R", "0.9"], {"min_identity": 0.93, "min_residues": 0.9}, 0.06, 0.4)])
def test_single_end_sam_identical_to_ngm(extra, filters, sub_rate, indels):
    """The second case tightens the output filters of GenericReadWriter::WriteRead (`-i`, `-R`): 17 % of the reads come out unmapped."""
    from nextgenmap_b200.host import pipeline
    with tempfile.TemporaryDirectory(prefix="se_") as td:
        d = Path(td)
        e2e.write_inputs(d, ref_len=500_000, n_reads=1500, read_len=100, seed=99, sub_rate=sub_rate, indel_reads=indels)
        want = [ln for ln in e2e.run("ref", d, threads=2, extra=["-s", "0.5", *extra]) if not ln.startswith("@")]
        orig = pipeline.sam_lines
        pipeline.sam_lines = lambda *a, **k: orig(*a, **filters, **k)
        try:
            got, _ = oracle_sam(d, 100, 0, 0.5, False)
        finally:
            pipeline.sam_lines = orig
    diff(got, want)
    if filters:
        assert sum(1 for ln in want if ln.split("\t")[1] == "4") > 100


def with_heap(batch):
    """The oracle batch in the layout ngm_b200_align_pairs returns: ALIGN_REC records + one string heap (CIGAR then MD per read)."""
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC
    n = len(batch.best_pair)
    recs = np.zeros(n, dtype=ALIGN_REC)
    heap = bytearray()
    for r in range(n):
        for f in ("position_offset", "qstart", "qend", "nm", "identity", "score"):
            recs[r][f] = batch.recs[r][f]
        if batch.best_pair[r] >= 0:
            cig, md = batch.strings(r)
            recs[r]["str_off"], recs[r]["cigar_len"], recs[r]["md_len"] = len(heap), len(cig), len(md)
            heap += cig + md
    batch.recs = recs
    batch.heap = np.frombuffer(bytes(heap) + b"\0", np.uint8)
    return batch


@pytest.mark.parametrize("paired", [False, True])
def test_native_sam_formatter_equals_host_mirror(paired):
    """ngm_b200_format_sam (C++, multi-threaded; host only) against the Python mirror that is checked against NGM's own SAM above."""
    from nextgenmap_b200.host import pipeline
    with tempfile.TemporaryDirectory(prefix="fmt_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=300_000, n_frags=600, read_len=100, seed=77)
        else:
            e2e.write_inputs(d, ref_len=300_000, n_reads=1200, read_len=100, seed=78, indel_reads=0.2)
        ref = LaidOutReference(d / "ref.fa")
        names, seqs, quals = read_fastq(d / "reads.fq", paired)
    reads = rows(seqs, 102)
    ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13)
    batch = with_heap(mapper_port.map_batch(ref.packed, ref.concat_len, ix, reads, 102, 20, 0, 0.5, mapper_port.Selector(), paired=paired))
    ix.close()
    want = pipeline.sam_lines_paired(batch, reads, names, quals, ref, 20) if paired else pipeline.sam_lines(None, batch, reads, names, quals, ref, 20)
    for threads in (1, 3):
        got = pipeline.format_sam(batch, reads, names, quals, ref.as_encoded_reference(), paired, threads=threads).decode().splitlines()
        assert got == want
    if paired:
        assert sorted(want) == gzip.open(GOLD / "pe_l100.sam.gz", "rt").read().splitlines()


def oracle_topn_sam(d: Path, read_len: int, mode: int, topn: int, strata: int):
    from nextgenmap_b200.host import EncodedReference, pipeline
    enc = d / "ref.fa-enc.2.ngm"
    ref = EncodedReference(str(enc)) if enc.exists() else LaidOutReference(d / "ref.fa")
    names, seqs, quals = read_fastq(d / "reads.fq", False)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = rows(seqs, qml)
    ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13)
    batch = mapper_port.map_batch_topn(ref.packed, ref.concat_len, ix, reads, qml, cor, mode, 0.5, topn, mapper_port.Selector(strata=strata))
    ix.close()
    got = sorted(pipeline.sam_lines_topn(batch, reads, names, quals, ref, cor))
    ref.close()
    return got


def test_topn_golden_sam():
    """`-n 3` (ScoreBuffer::topNSE, several alignments per read, 0x100 lines, the duplicate filter of GenericReadWriter::WriteRead):
    oracle run against the committed SAM of the unmodified NGM (tests/golden/make_pe_golden.py)."""
    with tempfile.TemporaryDirectory(prefix="topngold_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=300_000, n_frags=500, read_len=100, seed=79)
        got = oracle_topn_sam(d, 100, 0, 3, 0)
    want = gzip.open(GOLD / "se_topn3_l100.sam.gz", "rt").read().splitlines()
    diff(got, want)
    assert sum(1 for ln in want if int(ln.split("\t")[1]) & 0x100) > 100


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("extra,topn,strata,seed", [(["-n", "2", "--strata"], 2, 1, 42), (["-n", "5", "-e"], 5, 0, 43)])
def test_topn_sam_identical_to_ngm(extra, topn, strata, seed):
    with tempfile.TemporaryDirectory(prefix="topn_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=400_000, n_frags=600, read_len=100, seed=seed)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["-s", "0.5", *extra]) if not ln.startswith("@")]
        got = oracle_topn_sam(d, 100, 1 if "-e" in extra else 0, topn, strata)
    diff(got, want)


def test_native_sam_formatter_topn_equals_host_mirror():
    """ngm_b200_format_sam with several alignments per read (topn 3) against the Python mirror that is checked against `ngm -n 3`."""
    from nextgenmap_b200.host import pipeline
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC
    with tempfile.TemporaryDirectory(prefix="fmt_topn_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=300_000, n_frags=500, read_len=100, seed=79)
        ref = LaidOutReference(d / "ref.fa")
        names, seqs, quals = read_fastq(d / "reads.fq", False)
    reads = rows(seqs, 102)
    ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13)
    batch = mapper_port.map_batch_topn(ref.packed, ref.concat_len, ix, reads, 102, 20, 0, 0.5, 3, mapper_port.Selector())
    ix.close()
    want = pipeline.sam_lines_topn(batch, reads, names, quals, ref, 20)
    n, topn = batch.sel.shape
    recs, heap = np.zeros((n, topn), dtype=ALIGN_REC), bytearray()
    for r in range(n):
        for j in range(int(batch.n_sel[r])):
            for f in ("position_offset", "qstart", "qend", "nm", "identity", "score"):
                recs[r, j][f] = batch.recs[r, j][f]
            cig, md = batch.strings(r, j)
            recs[r, j]["str_off"], recs[r, j]["cigar_len"], recs[r, j]["md_len"] = len(heap), len(cig), len(md)
            heap += cig + md
    batch.recs, batch.heap = recs, np.frombuffer(bytes(heap) + b"\0", np.uint8)
    for threads in (1, 3):
        got = pipeline.format_sam(batch, reads, names, quals, ref.as_encoded_reference(), False, threads=threads).decode().splitlines()
        assert got == want
    assert sorted(want) == gzip.open(GOLD / "se_topn3_l100.sam.gz", "rt").read().splitlines()


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("paired", [False, True])
def test_min_mq_identical_to_ngm(paired):
    """`-Q 20`: reads below the mapping quality are written as unmapped without being aligned (AlignmentBuffer.cpp:46-49), for pairs
    through WritePair's mapped1 / mapped2 (GenericReadWriter.h:281-284); mirror and native formatter against NGM's SAM."""
    from nextgenmap_b200.host import EncodedReference, pipeline
    with tempfile.TemporaryDirectory(prefix="minmq_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=400_000, n_frags=700, read_len=100, seed=81)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["-s", "0.5", "-Q", "20"] + (["-p"] if paired else [])) if not ln.startswith("@")]
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq(d / "reads.fq", paired)
    reads = rows(seqs, 102)
    ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13)
    batch = with_heap(mapper_port.map_batch(ref.packed, ref.concat_len, ix, reads, 102, 20, 0, 0.5, mapper_port.Selector(), paired=paired))
    ix.close()
    got = pipeline.sam_lines_paired(batch, reads, names, quals, ref, 20, min_mq=20) if paired else pipeline.sam_lines(None, batch, reads, names, quals, ref, 20, min_mq=20)
    assert pipeline.format_sam(batch, reads, names, quals, ref, paired, min_mq=20).decode().splitlines() == got
    diff(sorted(got), want)
    assert sum(1 for ln in want if int(ln.split("\t")[1]) & 4) > 300
    ref.close()


def bisulfite(path, seed, paired):
    """Directional bisulfite chemistry on the reads of a FASTQ file: ~90 % of the C of a first mate / single read become T, ~90 % of the
    G of a second mate become A (what CS::PrefixMutateSearch undoes: T -> C, second mate A -> G; CS.cpp:362-380)."""
    lines = Path(path).read_bytes().split(b"\n")
    rng = np.random.default_rng(seed)
    for i in range(1, len(lines), 4):
        s = np.frombuffer(lines[i], np.uint8).copy()
        second = paired and ((i // 4) & 1)
        m = (s == ord("G" if second else "C")) & (rng.random(len(s)) < 0.9)
        s[m] = ord("A" if second else "T")
        lines[i] = s.tobytes()
    Path(path).write_bytes(b"\n".join(lines))


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("paired,seed", [(False, 71), (True, 72)])
def test_bs_mapping_sam_identical_to_ngm(paired, seed):
    """`ngm --bs-mapping` against the chain of restatements: mutated candidate search on an every-position index, the bs scoring scheme with the
    per-candidate direction flag, X-op accounting under bs_mapping, ZS:Z -- pins the direction rule for second mates and the SAM tag."""
    with tempfile.TemporaryDirectory(prefix="bs_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=300_000, n_frags=600, read_len=100, seed=seed)
        else:
            e2e.write_inputs(d, ref_len=400_000, n_reads=1_500, read_len=100, seed=seed, indel_reads=0.15)
        bisulfite(d / "reads.fq", seed, paired)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["--bs-mapping", "-s", "0.5"] + (["-p"] if paired else [])) if not ln.startswith("@")]
        sc = port.Scoring(match=4, mismatch=2, gap_read=10, gap_ref=10, bs_mapping=1, match_tt=4, match_tc=4)      # Config.cpp:463-469
        got, _ = oracle_sam(d, 100, 0, 0.5, paired, mutate={"mode": 1, "bs_cutoff": 6, "read_skip": 2}, scoring=sc)
    diff(got, want)
    zs = {f for ln in want for f in ln.split("\t")[11:] if f.startswith("ZS:Z:")}
    assert zs == ({"ZS:Z:++", "ZS:Z:-+", "ZS:Z:--", "ZS:Z:+-"} if paired else {"ZS:Z:++", "ZS:Z:-+"})
    assert sum(1 for ln in want if not int(ln.split("\t")[1]) & 4) > 0.8 * len(want)


def slam_convert(path, seed, paired):
    """SLAMseq chemistry: ~6 % of the T of a first mate / single read are read as C, ~6 % of the A of a second mate as G."""
    lines = Path(path).read_bytes().split(b"\n")
    rng = np.random.default_rng(seed)
    for i in range(1, len(lines), 4):
        s = np.frombuffer(lines[i], np.uint8).copy()
        second = paired and ((i // 4) & 1)
        m = (s == ord("A" if second else "T")) & (rng.random(len(s)) < 0.06)
        s[m] = ord("G" if second else "C")
        lines[i] = s.tobytes()
    Path(path).write_bytes(b"\n".join(lines))


@pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")
@pytest.mark.parametrize("paired,slam,seed,estimate", [(False, 6, 81, False), (True, 7, 82, False), (False, 1, 83, False), (False, 6, 84, True)])
def test_slam_seq_sam_identical_to_ngm(paired, slam, seed, estimate):
    """`ngm --slam-seq <bits>`: bit 1 (2) the T>C tolerant scoring scheme with the per-candidate direction flag, bit 2 (4) the weighted k-mer
    mutation in candidate search, any bit the per-column record behind the TC:i / RA:Z / MP:Z tags (Align::ExtendedData)."""
    with tempfile.TemporaryDirectory(prefix="slam_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=300_000, n_frags=500, read_len=100, seed=seed)
        else:
            e2e.write_inputs(d, ref_len=400_000, n_reads=1_200, read_len=100, seed=seed, indel_reads=0.2)
        slam_convert(d / "reads.fq", seed, paired)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["--slam-seq", str(slam)] + ([] if estimate else ["-s", "0.5"]) + (["-p"] if paired else []))
                if not ln.startswith("@")]
        sens = 0.5
        if estimate:                                             # no -s: ReadProvider::init estimates with plain k-mers also under slam_seq
            ref = LaidOutReference(d / "ref.fa")
            ix = cs_port.Index(ref.packed, ref.concat_len, [(c[1], c[2]) for c in ref.contigs], k=13)
            sens, _ = ix.estimate_sensitivity(rows(read_fastq(d / "reads.fq", paired)[1], 102))
            ix.close()
            assert "%f" % sens == "%f" % e2e.logged_sensitivity() and 0.3 < sens < 0.9
        sc = port.Scoring(slam_seq=slam, match_tt=10, match_tc=2)                                     # Config.cpp:446-447
        got, _ = oracle_sam(d, 100, 0, sens, paired, mutate={"mode": 2} if slam & 4 else None, scoring=sc)
    diff(got, want)
    assert sum(1 for ln in want if "\tMP:Z:" in ln) > 0.5 * len(want) and sum(1 for ln in want if "\tTC:i:0" not in ln and "\tTC:i:" in ln) > 0.3 * len(want)
