"""The whole single-end mapping run on the device against the UNMODIFIED NextGenMap (oracle/_ref/ngm/ngm_ref: its own CS, its own
OpenCL kernels on the CPU, its own selection and SAM writer): BASELINE configs[0]-shaped input (100 bp reads vs a 5 Mbp contig), and a
two-contig / indel-rich / 150 bp variant.  The device side: prefix table built on the device from the reference file NGM wrote, k-mer vote
(cs_search), BatchScore of every candidate, top-1 + MAPQ + NH, BatchAlign, rendered by the host mirror of SAMWriter.  Sorted SAM bodies
must be byte-identical: POS, FLAG, MAPQ, CIGAR, AS, NM, NH/X0, XI, XE (= the best k-mer vote), XR, MD -- for every read."""
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import cs_port, ngm_e2e as e2e

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")]


def read_fastq(path):
    names, seqs, quals = [], [], []
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:].decode().split()[0])
        seqs.append(lines[i + 1])
        quals.append(lines[i + 3])
    return names, seqs, quals


# the last three: BASELINE configs[3] (250 bp, 12 % substitutions + 1.5 % insertions + 1.5 % deletions per base, `-C 40` => corridor 80: the
# wide-band forward pass and the many-candidates regime of candidate search), the same in end-free mode, and the 400 bp point of configs[4]
@pytest.mark.parametrize("ref_len,n_reads,read_len,extra", [(5_000_000, 10_000, 100, []), (1_200_000, 4_000, 150, ["-e"]), (800_000, 3_000, 75, ["-s", "0.8"]),
                                                             (900_000, 5_000, 120, ["--estimate"]), (1_500_000, 3_000, 250, ["-C", "40", "--divergent"]),
                                                             (700_000, 1_500, 250, ["-C", "40", "--divergent", "-e"]), (1_000_000, 2_000, 400, [])])
def test_sam_identical_to_ngm(ref_len, n_reads, read_len, extra):
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    divergent = "--divergent" in extra
    extra = [a for a in extra if a != "--divergent"]
    estimate = "--estimate" in extra                # no -s: NGM estimates the sensitivity from every 1000th read (ReadProvider.cpp:236-325)
    sens = float(extra[extra.index("-s") + 1]) if "-s" in extra else 0.5
    mode = 1 if "-e" in extra else 0
    with tempfile.TemporaryDirectory(prefix="pipe_") as td:
        d = Path(td)
        if divergent:
            e2e.write_inputs(d, ref_len=ref_len, n_reads=n_reads, read_len=read_len, seed=4242 + read_len + mode, sub_rate=0.12, indel_reads=0.0, ins_rate=0.015,
                             del_rate=0.015)
        else:
            e2e.write_inputs(d, ref_len=ref_len, n_reads=n_reads, read_len=read_len, seed=4242 + read_len, indel_reads=0.15)
        args = [a for a in extra if a not in ("-s", str(sens), "--estimate")] + ([] if estimate else ["-s", str(sens)])
        want = [ln for ln in e2e.run("ref", d, threads=4, extra=args) if not ln.startswith("@")]
        logged = e2e.logged_sensitivity() if estimate else None
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq(d / "reads.fq")
    qml = (read_len | 1) + 1                         # ReadProvider.cpp:288
    cor = 2 * int(extra[extra.index("-C") + 1]) if "-C" in extra else int(5 + 0.15 * read_len)      # Config.cpp:540-557, ReadProvider.cpp:304
    reads = np.zeros((len(seqs), qml), np.uint8)
    for i, s in enumerate(seqs):
        reads[i, : len(s)] = np.frombuffer(s, np.uint8)
    sw = CudaSW(qml, cor)
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, sensitivity=sens))
    if estimate:
        got_sens = sw.cs_estimate_sensitivity(reads)            # installs it for the searches below
        assert "%f" % got_sens == "%f" % logged and 0.3 < got_sens < 0.9
    batch = pipeline.map_reads(sw, reads, mode)
    got = sorted(pipeline.sam_lines(sw, batch, reads, names, quals, ref, cor))
    native = sorted(pipeline.format_sam(batch, reads, names, quals, ref, False).decode().splitlines())      # ngm_b200_format_sam (C++, threads)
    assert native == got
    one_call = pipeline.map_batch(sw, reads, mode)                 # ngm_b200_map_batch: the whole batch in one C call
    assert sorted(pipeline.format_sam(one_call, reads, names, quals, ref, False).decode().splitlines()) == got
    # the same candidates through ngm_b200_run_batch (packed reads, 64-bit descriptors, several lanes, single-candidate reads fused)
    sw.set_pipeline(3, 1024)
    rb = sw.run_batch(mode, reads, batch.cand_begin, batch.pairs, packed=True, desc_u64=True)
    assert np.array_equal(rb["best_pair"], batch.best_pair) and np.array_equal(rb["mapq"], batch.mapq) and np.array_equal(rb["num_top"], batch.num_top)
    assert np.array_equal(rb["scores"].view(np.uint32), batch.scores.view(np.uint32))
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        assert np.array_equal(rb["recs"][f], batch.recs[f]), f
    assert all(sw.strings_of(rb["recs"], rb["heap"], r) == batch.strings(r) for r in range(len(reads)) if batch.recs[r]["score"] >= 0)
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    assert sum(1 for ln in want if ln.split("\t")[1] != "4") > (0.8 if divergent else 0.95) * len(want)
    sw.close()
    ref.close()


@pytest.mark.parametrize("ref_len,n_frags,read_len,seed,extra", [(600_000, 2000, 100, 31, []), (500_000, 1200, 150, 32, ["-e"]),
                                                                 (400_000, 1000, 75, 33, ["--fast-pairing"])])
def test_paired_sam_identical_to_ngm(ref_len, n_frags, read_len, seed, extra):
    """BASELINE configs[2] shape: interleaved mates, `ngm -p -t 1` (one CS thread: the insert-size mean that breaks ties between equally
    scoring pairs runs through the reads in input order).  Device: cs_search + BatchScore per mate, ngm_b200_dev_select_pairs (top1PE),
    BatchAlign; pair check, filters and SAM flags / RNEXT / PNEXT / TLEN by the host mirror of WriteRead / WritePair / DoWritePair.  The
    reads go through in three batches."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    from tests.test_mapper_oracle import read_fastq as read_fastq_pe, rows
    mode = 1 if "-e" in extra else 0
    with tempfile.TemporaryDirectory(prefix="pipe_pe_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=ref_len, n_frags=n_frags, read_len=read_len, seed=seed)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["-p", "-s", "0.5", *extra]) if not ln.startswith("@")]
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq_pe(d / "reads.fq", True)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = rows(seqs, qml)
    sw = CudaSW(qml, cor)
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, sensitivity=0.5))
    sw.pe_configure(fast_pairing=1 if "--fast-pairing" in extra else 0)
    got, native, n = [], [], len(names)
    step = (n // 3 + 1) & ~1
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        batch = pipeline.map_pairs(sw, reads[lo:hi], mode)
        got += pipeline.sam_lines_paired(batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, cor)
        native += pipeline.format_sam(batch, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, True).decode().splitlines()
    assert native == got                                # ngm_b200_format_sam (C++, threads): same lines, same order
    sw.pe_configure(fast_pairing=1 if "--fast-pairing" in extra else 0)        # again from the start, through ngm_b200_map_batch
    one_call = []
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        # undersized arrays first: the call reports what it needs and leaves the insert-size sums as they were
        mb = pipeline.map_batch(sw, reads[lo:hi], mode, paired=True, capacity=64 if lo else 0, heap_bytes=512)
        one_call += pipeline.format_sam(mb, reads[lo:hi], names[lo:hi], quals[lo:hi], ref, True).decode().splitlines()
    assert one_call == got
    got.sort()
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    flags = {ln.split("\t")[1] for ln in want}
    assert {"99", "147", "83", "163"} <= flags and len(flags) >= 10
    sw.close()
    ref.close()


@pytest.mark.parametrize("extra,topn,strata,seed", [(["-n", "3"], 3, False, 51), (["-n", "2", "--strata"], 2, True, 52)])
def test_topn_sam_identical_to_ngm(extra, topn, strata, seed):
    """`ngm -n <topn>` (ScoreBuffer::topNSE): several alignments per read, secondary lines (0x100), repeated locations dropped."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    from tests.test_mapper_oracle import read_fastq as read_fastq_pe, rows
    with tempfile.TemporaryDirectory(prefix="pipe_topn_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=500_000, n_frags=900, read_len=100, seed=seed)      # repeats: several equally good candidates
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["-s", "0.5", *extra]) if not ln.startswith("@")]
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq_pe(d / "reads.fq", False)
    reads = rows(seqs, 102)
    sw = CudaSW(102, 20)
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, sensitivity=0.5))
    batch = pipeline.map_reads_topn(sw, reads, topn, strata)
    got = sorted(pipeline.sam_lines_topn(batch, reads, names, quals, ref, 20))
    assert sorted(pipeline.format_sam(batch, reads, names, quals, ref, False).decode().splitlines()) == got      # ngm_b200_format_sam, topn records
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    assert sum(1 for ln in want if int(ln.split("\t")[1]) & 0x100) > (5 if strata else 100)
    # the same run through the one-call entry points (ngm_b200_se_configure_topn): several sub-batches over the lanes
    sw.se_configure(strata=1 if strata else 0, topn=topn)
    sw.set_pipeline(3, 256)
    one_call = pipeline.map_batch(sw, reads)
    assert np.array_equal(one_call.sel, batch.sel) and np.array_equal(one_call.n_sel, batch.n_sel) and np.array_equal(one_call.best_pair, batch.sel[:, 0])
    assert sorted(pipeline.format_sam(one_call, reads, names, quals, ref, False).decode().splitlines()) == got
    rb = sw.run_batch(0, reads, batch.cand_begin, batch.pairs, packed=True, desc_u64=True)
    assert np.array_equal(rb["sel"], batch.sel) and np.array_equal(rb["n_sel"], batch.n_sel) and np.array_equal(rb["best_pair"], batch.sel[:, 0])
    assert np.array_equal(rb["mapq"], batch.mapq) and np.array_equal(rb["num_top"], batch.num_top)
    assert np.array_equal(rb["scores"].view(np.uint32), batch.scores.view(np.uint32))
    taken = batch.sel >= 0
    assert np.all(rb["recs"]["score"][~taken] == -1.0)
    for f in ("position_offset", "qstart", "qend", "nm", "score", "cigar_len", "md_len"):
        assert np.array_equal(rb["recs"][f][taken], batch.recs[f][taken]), f
    for r in range(len(reads)):
        for j in range(int(batch.n_sel[r])):
            assert sw.strings_of(rb["recs"][r], rb["heap"], j) == batch.strings(r, j)
    sw.se_configure(0, 1)
    sw.close()
    ref.close()


@pytest.mark.parametrize("flag,paired", [("--hard-clip", False), ("--silent-clip", False), ("--hard-clip", True)])
def test_clipping_and_read_group_identical_to_ngm(flag, paired):
    """`--hard-clip` / `--silent-clip` (H ops or none in the CIGAR, SEQ / QUAL without the clipped ends) and `--rg-id` (RG:Z on every record,
    mapped or not): SAMWriter.cpp:104,146-168,358-360 -- the library's formatter (ngm_b200_sam_opts.clip_seq / read_group) against ngm."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    from tests.test_mapper_oracle import read_fastq as read_fastq_pe, rows
    hard = flag == "--hard-clip"
    with tempfile.TemporaryDirectory(prefix="pipe_clip_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=500_000, n_frags=1000, read_len=100, seed=77)
        else:
            e2e.write_inputs(d, ref_len=900_000, n_reads=3000, read_len=150, seed=78, sub_rate=0.06, indel_reads=0.2)
        lines = (d / "reads.fq").read_bytes().split(b"\n")          # position-dependent qualities: clipping QUAL at the wrong end must show
        for i in range(3, len(lines), 4):
            lines[i] = bytes(40 + (i + 3 * k) % 33 for k in range(len(lines[i])))
        (d / "reads.fq").write_bytes(b"\n".join(lines))
        args = ["-s", "0.5", flag, "--rg-id", "lane7"] + (["-p"] if paired else [])
        want = [ln for ln in e2e.run("ref", d, threads=1 if paired else 4, extra=args) if not ln.startswith("@")]
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        if paired:
            names, seqs, quals = read_fastq_pe(d / "reads.fq", True)
        else:
            names, seqs, quals = read_fastq(d / "reads.fq")
    read_len = max(len(x) for x in seqs)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = np.zeros((len(seqs), qml), np.uint8)
    for i, x in enumerate(seqs):
        reads[i, : len(x)] = np.frombuffer(x, np.uint8)
    sw = CudaSW(qml, cor, hard_clip=1 if hard else 0, silent_clip=0 if hard else 1)
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, sensitivity=0.5))
    if paired:
        sw.pe_configure()
    batch = pipeline.map_batch(sw, reads, 0, paired=paired)
    got = pipeline.format_sam(batch, reads, names, quals, ref, paired, clip_seq=True, read_group="lane7").decode().splitlines()
    got.sort()
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    assert all("RG:Z:lane7" in ln for ln in want)
    clipped = [ln for ln in want if ln.split("\t")[1] != "4" and len(ln.split("\t")[9]) < read_len]
    assert len(clipped) > 20                                  # reads that really lost bases
    if hard:
        assert any("H" in ln.split("\t")[5] for ln in clipped)
    else:
        assert not any("H" in ln.split("\t")[5] or "S" in ln.split("\t")[5] for ln in want)
    sw.close()
    ref.close()


def _bisulfite(path, seed, paired):
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


@pytest.mark.parametrize("paired,seed", [(False, 71), (True, 72)])
def test_bs_mapping_sam_identical_to_ngm(paired, seed):
    """`ngm --bs-mapping` (single-end and `-p`): candidate search with mutated k-mers on an every-position index (CS::PrefixMutateSearch),
    the bs scoring scheme (match 4, mismatch 2, gaps 10 / 10, T-C pairs by direction; Config.cpp:463-469, SWOcl.cpp:228-242) with the
    direction flag per candidate, X-op accounting of computeCigarMD under bs_mapping, ZS:Z in the SAM record (SAMWriter.cpp:173-187)."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    from tests.test_mapper_oracle import read_fastq as read_fastq_pe, rows
    read_len = 100
    with tempfile.TemporaryDirectory(prefix="pipe_bs_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=300_000, n_frags=600, read_len=read_len, seed=seed)
        else:
            e2e.write_inputs(d, ref_len=400_000, n_reads=1_500, read_len=read_len, seed=seed, indel_reads=0.15)
        _bisulfite(d / "reads.fq", seed, paired)
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["--bs-mapping", "-s", "0.5"] + (["-p"] if paired else [])) if not ln.startswith("@")]
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq_pe(d / "reads.fq", paired)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = rows(seqs, qml)
    sw = CudaSW(qml, cor, match_bonus=4, mismatch_penalty=2, gap_read_penalty=10, gap_ref_penalty=10, match_bonus_tt=4, match_bonus_tc=4, bs_mapping=1)
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, kmer_skip=0, sensitivity=0.5))
    sw.cs_configure_mutation(bs_mapping=1, bs_cutoff=6, paired=paired, read_kmer_skip=2)
    if paired:
        sw.pe_configure()
        batch = pipeline.map_pairs(sw, reads, 0)
        batch.bs_mapping = 1
        got = pipeline.sam_lines_paired(batch, reads, names, quals, ref, cor)
    else:
        batch = pipeline.map_reads(sw, reads, 0)
        batch.bs_mapping = 1
        got = pipeline.sam_lines(sw, batch, reads, names, quals, ref, cor)
    native = pipeline.format_sam(batch, reads, names, quals, ref, paired, bs_mapping=1).decode().splitlines()
    assert native == got
    if paired:
        sw.pe_configure()
    sw.set_pipeline(3, 256)                                        # several lanes, sub-batches of whole pairs
    one_call = pipeline.map_batch(sw, reads, 0, paired=paired)
    assert pipeline.format_sam(one_call, reads, names, quals, ref, paired, bs_mapping=1).decode().splitlines() == got
    got.sort()
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    zs = {f for ln in want for f in ln.split("\t")[11:] if f.startswith("ZS:Z:")}
    assert zs == ({"ZS:Z:++", "ZS:Z:-+", "ZS:Z:--", "ZS:Z:+-"} if paired else {"ZS:Z:++", "ZS:Z:-+"})
    assert sum(1 for ln in want if not int(ln.split("\t")[1]) & 4) > 0.8 * len(want)
    sw.close()
    ref.close()


@pytest.mark.parametrize("paired,slam,seed,estimate", [(False, 6, 81, False), (True, 7, 82, False), (False, 6, 84, True)])
def test_slam_seq_sam_identical_to_ngm(paired, slam, seed, estimate):
    """`ngm --slam-seq <bits>`: weighted k-mer mutation in candidate search (bit 2: fractional votes, XE:i is their integer part), the T>C
    tolerant scoring scheme with the direction flag (bit 1), TC:i / RA:Z / MP:Z in the SAM record from CIGAR + MD + read."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host import pipeline
    from tests.test_mapper_oracle import read_fastq as read_fastq_pe, rows, slam_convert
    read_len = 100
    with tempfile.TemporaryDirectory(prefix="pipe_slam_") as td:
        d = Path(td)
        if paired:
            e2e.write_paired_inputs(d, ref_len=300_000, n_frags=500, read_len=read_len, seed=seed)
        else:
            e2e.write_inputs(d, ref_len=400_000, n_reads=1_200, read_len=read_len, seed=seed, indel_reads=0.2)
        slam_convert(d / "reads.fq", seed, paired)
        # estimate: no -s -- NGM estimates the sensitivity with PLAIN k-mers (ReadProvider::init), the run itself mutates them
        want = [ln for ln in e2e.run("ref", d, threads=1, extra=["--slam-seq", str(slam)] + ([] if estimate else ["-s", "0.5"]) + (["-p"] if paired else []))
                if not ln.startswith("@")]
        logged = e2e.logged_sensitivity() if estimate else None
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
        names, seqs, quals = read_fastq_pe(d / "reads.fq", paired)
    qml, cor = (read_len | 1) + 1, int(5 + 0.15 * read_len)
    reads = rows(seqs, qml)
    sw = CudaSW(qml, cor, match_bonus_tt=10, match_bonus_tc=2, slam_seq=slam)                       # Config.cpp:446-447
    sw.set_reference(ref.packed, ref.concat_len)
    sw.cs_build_index([(c[1], c[2]) for c in ref.contigs], sw.cs_params(kmer=13, sensitivity=0.5))
    sw.cs_configure_mutation(slam_seq=slam, paired=paired)
    if estimate:
        got_sens = sw.cs_estimate_sensitivity(reads)               # after the mutation was switched on: it must not take part
        assert "%f" % got_sens == "%f" % logged and 0.3 < got_sens < 0.9
    if paired:
        sw.pe_configure()
        batch = pipeline.map_pairs(sw, reads, 0)
        batch.slam_seq = slam
        got = pipeline.sam_lines_paired(batch, reads, names, quals, ref, cor)
    else:
        batch = pipeline.map_reads(sw, reads, 0)
        batch.slam_seq = slam
        got = pipeline.sam_lines(sw, batch, reads, names, quals, ref, cor)
    assert pipeline.format_sam(batch, reads, names, quals, ref, paired, slam_seq=slam).decode().splitlines() == got
    if paired:
        sw.pe_configure()
    sw.set_pipeline(3, 256)
    one_call = pipeline.map_batch(sw, reads, 0, paired=paired)
    assert pipeline.format_sam(one_call, reads, names, quals, ref, paired, slam_seq=slam).decode().splitlines() == got
    got.sort()
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, f"{len(bad)} of {len(want)} SAM lines differ, first:\n{bad[0][0]}\n{bad[0][1]}"
    assert sum(1 for ln in want if "\tMP:Z:" in ln) > 0.5 * len(want)
    sw.close()
    ref.close()
