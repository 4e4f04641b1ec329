"""NGM's encoded-reference cache file `<ref>-enc.2.ngm`: our reader against a file written by the
unmodified NextGenMap itself (oracle/_ref/ngm/ngm_ref, `ngm -r ref.fa` = pre-process only), and convert()
(SequenceProvider.cpp:111-141) against the layout rules of SequenceProvider::Init (1000 N in front of / between /
after the contigs, SequenceProvider.cpp:289-330)."""
import os
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import ngm_e2e as e2e, port

pytestmark = pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")


def write_fasta(path, contigs):
    with open(path, "wb") as f:
        for name, seq in contigs:
            f.write(b">" + name + b"\n")
            for i in range(0, len(seq), 70):
                f.write(seq[i: i + 70] + b"\n")


def test_reader_matches_file_written_by_ngm():
    from nextgenmap_b200.host import EncodedReference
    rng = np.random.default_rng(11)
    mk = lambda n: bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)])
    contigs = [(b"chrA first", mk(40_001)), (b"chrB", mk(25_000)[:12_000] + b"N" * 37 + mk(13_000)), (b"chrC", mk(9_999))]
    with tempfile.TemporaryDirectory(prefix="encref_") as td:
        d = Path(td)
        write_fasta(d / "ref.fa", contigs)
        env = dict(os.environ)
        ocl = e2e.HERE / "_ref" / "ocl"
        env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
        env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        p = subprocess.run([str(e2e.binary("ref")), "-r", str(d / "ref.fa"), "--no-progress"], env=env, capture_output=True, text=True, cwd=d)
        enc = d / "ref.fa-enc.2.ngm"
        assert enc.exists(), p.stdout + p.stderr
        ref = EncodedReference(str(enc))
        ref.write(str(d / "copy-enc.2.ngm"))                          # ngm_b200_write_enc_ref: the same bytes NGM wrote
        assert (d / "copy-enc.2.ngm").read_bytes() == enc.read_bytes()
    # layout: 1000 N, contig (+1 N if odd), 1000 N, ...
    expect = b"N" * 1000
    starts = []
    for _, seq in contigs:
        starts.append(len(expect))
        expect += seq + (b"N" if len(seq) & 1 else b"") + b"N" * 1000
    assert [c[0] for c in ref.contigs] == ["chrA", "chrB", "chrC"]
    assert [c[1] for c in ref.contigs] == starts
    assert [c[2] for c in ref.contigs] == [len(s) for _, s in contigs]
    assert ref.concat_len == len(expect) - 1                          # GetConcatRefLen() = binRefIndex - 1
    mine = port.pack_ref(expect)[: len(expect) // 2]
    np.testing.assert_array_equal(ref.packed[: len(mine)], mine)      # same 4-bit packing, high nibble first
    # convert(): inside contigs, in spacers, at the edges
    assert ref.convert(starts[0]) == (0, 0)
    assert ref.convert(starts[1] + 12_345) == (1, 12_345)
    assert ref.convert(starts[2] + 9_998) == (2, 9_998)
    assert ref.convert(starts[1] - 1) is None and ref.convert(starts[1] - 999) is None      # < 1000 in front of the next start
    assert ref.convert(starts[1] - 1000) == (0, starts[1] - 1000 - starts[0])              # exactly 1000 away still counts as contig 0
    assert ref.convert(5) is None                                                          # leading spacer
    ref.close()


@pytest.mark.gpu
def test_descriptor_path_on_a_reference_file_written_by_ngm():
    """Windows fetched from the file's packing score exactly like windows decoded by the oracle from the same bytes."""
    from nextgenmap_b200.host import CudaSW, EncodedReference
    from nextgenmap_b200.host.cuda_sw import PAIR
    rng = np.random.default_rng(12)
    seq = bytes(np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, 300_000)])
    with tempfile.TemporaryDirectory(prefix="encref_") as td:
        d = Path(td)
        write_fasta(d / "ref.fa", [(b"chr1", seq)])
        env = dict(os.environ)
        ocl = e2e.HERE / "_ref" / "ocl"
        env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
        env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        subprocess.run([str(e2e.binary("ref")), "-r", str(d / "ref.fa"), "--no-progress"], env=env, capture_output=True, cwd=d)
        ref = EncodedReference(str(d / "ref.fa-enc.2.ngm"))
    qml, cor, n = 102, 20, 2000
    sw = CudaSW(qml, cor)
    sw.set_reference(ref.packed, ref.concat_len)
    reads = np.zeros((n, qml), np.uint8)
    pairs = np.zeros(n, dtype=PAIR)
    refs = np.zeros((n, ((qml + cor) | 1) + 1), np.uint8)
    for i in range(n):
        pos = int(rng.integers(0, len(seq) - 110))
        s = np.frombuffer(seq[pos: pos + 100], np.uint8).copy()
        mut = rng.random(100) < 0.03
        s[mut] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(mut.sum()))]
        reads[i, :100] = s
        start = 1000 + pos + int(rng.integers(-2, 3)) - (cor >> 1)
        pairs[i] = (start, i, 0)
        refs[i] = np.frombuffer(port.decode_window(ref.packed, ref.concat_len, start, refs.shape[1]), np.uint8)
    sw.set_reads(reads)
    for mode in (0, 1):
        np.testing.assert_array_equal(sw.score_pairs(mode, pairs), port.batch_score(refs, reads, qml, cor, mode))
    sw.close()
    ref.close()
