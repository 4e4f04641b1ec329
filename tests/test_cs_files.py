"""NGM's prefix-table cache file `<ref>-ht-<k>-<skip>.3.ngm`: our host reader / writer against a file the
unmodified NextGenMap wrote (oracle/_ref/ngm/ngm_ref, `ngm -r ref.fa -k 10` = pre-process only)."""
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from oracle import cs_port, ngm_e2e as e2e
from tests import cs_cases

pytestmark = pytest.mark.skipif(not e2e.available("ref"), reason="oracle/_ref/ngm/ngm_ref not built")


def test_reader_and_writer_round_trip_a_file_written_by_ngm():
    from nextgenmap_b200.host import PrefixTableFile
    contigs = cs_cases.make_reference(31)
    with tempfile.TemporaryDirectory(prefix="htfile_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        p = subprocess.run([str(e2e.binary("ref")), "-r", "ref.fa", "-k", "10", "--no-progress"], env=cs_port.ngm_env(), capture_output=True, text=True, cwd=d)
        path = d / "ref.fa-ht-10-2.3.ngm"
        assert path.exists(), p.stdout + p.stderr
        want = cs_port.read_ht_file(path)                     # numpy restatement of the layout
        ht = PrefixTableFile(str(path))
        assert (ht.kmer, ht.kmer_skip, ht.index_len, ht.table_len, ht.unit_offset) == (10, 2, 4 ** 10 + 1, want["table_len"], 0)
        np.testing.assert_array_equal(ht.tab, want["tab"])
        np.testing.assert_array_equal(ht.weight, want["weight"])
        np.testing.assert_array_equal(ht.table, want["table"])
        PrefixTableFile.write(str(d / "copy.ngm"), ht.kmer, ht.kmer_skip, ht.tab, ht.weight, ht.table)
        assert (d / "copy.ngm").read_bytes() == path.read_bytes()      # byte-identical to what NGM wrote


def test_reader_rejects_garbage(tmp_path):
    from nextgenmap_b200.host import PrefixTableFile, NgmB200Error
    (tmp_path / "bad.ngm").write_bytes(b"\0" * 64)
    with pytest.raises(NgmB200Error):
        PrefixTableFile(str(tmp_path / "bad.ngm"))
