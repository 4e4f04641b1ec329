"""BASELINE configs[0] end to end: the UNMODIFIED NextGenMap, once with its own OpenCL backend (CPU device) and
once with the CUDA backend swapped in at link time (nextgenmap_b200/link_seam, zero source changes), must write
the same SAM: same positions, MAPQ, CIGAR, AS / NM / XI / MD tags for every read.  Two CUDA builds: `ngm_cuda` with the
re-plumbed ScoreBuffer / AlignmentBuffer (their window fetch submits descriptors against the reference resident in HBM, one
shared copy for all CS threads; link_seam/replumb_shim.cpp) and `ngm_cuda_strict` (char** windows through IAlignment as they are).

Both binaries are built by oracle/Makefile.ngm where /root/reference exists and travel to the GPU box in
oracle/_ref/ (git-ignored); the test is skipped when they are absent.
"""
import tempfile
from pathlib import Path

import pytest

from oracle import ngm_e2e as e2e

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (e2e.available("ref") and e2e.available("cuda")), reason="oracle/_ref/ngm not built")]


def compare(extra, n_reads=10_000, read_len=100, threads=4, chemistry=None, ref_len=5_000_000, paired=False):
    with tempfile.TemporaryDirectory(prefix="ngm_e2e_") as td:
        d = Path(td)
        if paired:                                         # n_reads mates = n_reads / 2 fragments; some fragments are unmappable by construction
            e2e.write_paired_inputs(d, ref_len=ref_len, n_frags=n_reads // 2, read_len=read_len, seed=9)
        else:
            e2e.write_inputs(d, ref_len=ref_len, n_reads=n_reads, read_len=read_len)
        if chemistry is not None:
            chemistry(d / "reads.fq", 5, paired)
        want = e2e.run("ref", d, threads=threads, extra=extra, out_name="ref.sam")
        got = e2e.run("cuda", d, threads=threads, extra=extra, out_name="cuda.sam")
        strict = e2e.run("cuda_strict", d, threads=threads, extra=extra, out_name="cuda_strict.sam") if e2e.available("cuda_strict") else got
    assert strict == got
    n_head = sum(1 for ln in want if ln.startswith("@"))
    assert len(want) == n_reads + n_head                 # @HD, @SQ ... + one record per read
    mapped = sum(1 for ln in want if not ln.startswith("@") and ln.split("\t")[2] != "*")
    assert mapped > (0.8 if paired else 0.98) * n_reads
    diff = [(a, b) for a, b in zip(want, got) if a != b]
    assert len(want) == len(got) and not diff, f"{len(diff)} SAM lines differ, first:\n{diff[0][0]}\n{diff[0][1]}"


def test_config0_local_mode_sam_identical():
    """10 k x 100 bp SE vs 5 Mbp, -t 4 (BASELINE.json configs[0])."""
    compare(extra=())


def test_config0_end_to_end_mode_sam_identical():
    compare(extra=("-e",), n_reads=4000)


def test_150bp_single_thread_sam_identical():
    compare(extra=(), n_reads=4000, read_len=150, threads=1)


def test_bs_mapping_sam_identical():
    """`--bs-mapping`: NGM's own mutated candidate search feeds the backend the bs scoring scheme with a direction flag per candidate
    (ScoreBuffer.cpp:92-110,127) -- through the descriptors of the re-plumbed callers and through the strict char** path."""
    from tests.test_mapper_oracle import bisulfite
    compare(extra=("--bs-mapping",), n_reads=3000, threads=2, chemistry=bisulfite, ref_len=1_000_000)


def test_slam_seq_sam_identical():
    """`--slam-seq 7`: T>C tolerant scoring, and BatchAlign leaves Align::ExtendedData (one AlignmentPosition per aligned column), from which
    NGM's SAMWriter prints TC:i / RA:Z / MP:Z."""
    from tests.test_mapper_oracle import slam_convert
    compare(extra=("--slam-seq", "7"), n_reads=3000, threads=2, chemistry=slam_convert, ref_len=1_000_000)


def test_bs_mapping_paired_sam_identical():
    """`-p --bs-mapping -t 1`: second mates mutate / score the complementary bases (CS.cpp:362-380) and carry the inverted direction flag
    (ScoreBuffer.cpp:98-106, AlignmentBuffer.cpp:84-94); one CS thread, so that the insert-size tie-break sees the pairs in input order."""
    from tests.test_mapper_oracle import bisulfite
    compare(extra=("--bs-mapping", "-p"), n_reads=2000, threads=1, chemistry=bisulfite, ref_len=600_000, paired=True)
