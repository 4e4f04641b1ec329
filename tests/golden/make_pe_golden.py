"""Generate tests/golden/pe_l100.sam.gz and tests/golden/se_topn3_l100.sam.gz: the sorted SAM body the UNMODIFIED NextGenMap writes for a seeded paired-end input
(oracle/ngm_e2e.write_paired_inputs: repeated segments, broken pairs, unmappable mates), run as `ngm -p -t 1 -s 0.5`.

    python __graft_entry__.py            # builds oracle/_ref/ngm/ngm_ref
    python tests/golden/make_pe_golden.py
"""
from __future__ import annotations

import gzip
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import ngm_e2e as e2e  # noqa: E402

if __name__ == "__main__":
    with tempfile.TemporaryDirectory(prefix="pegold_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=300_000, n_frags=600, read_len=100, seed=77)
        body = [ln for ln in e2e.run("ref", d, threads=1, extra=["-p", "-s", "0.5"]) if not ln.startswith("@")]
    with gzip.GzipFile(Path(__file__).resolve().parent / "pe_l100.sam.gz", "wb", mtime=0) as f:
        f.write(("\n".join(body) + "\n").encode())
    print(len(body), "lines")
    # single-end with topn 3 (ScoreBuffer::topNSE): the same fragments read as single-end input
    with tempfile.TemporaryDirectory(prefix="topngold_") as td:
        d = Path(td)
        e2e.write_paired_inputs(d, ref_len=300_000, n_frags=500, read_len=100, seed=79)
        body = [ln for ln in e2e.run("ref", d, threads=1, extra=["-n", "3", "-s", "0.5"]) if not ln.startswith("@")]
    with gzip.GzipFile(Path(__file__).resolve().parent / "se_topn3_l100.sam.gz", "wb", mtime=0) as f:
        f.write(("\n".join(body) + "\n").encode())
    print(len(body), "lines (topn 3)")
