"""Generate tests/golden/cs_sensitivity.json: the sensitivity the UNMODIFIED NextGenMap estimates (run without -s) for seeded inputs,
taken from its log line 'Estimated sensitivity: %f' (ReadProvider.cpp:357).

    python __graft_entry__.py            # builds oracle/_ref/ngm/ngm_ref
    python tests/golden/make_cs_sensitivity_golden.py
"""
from __future__ import annotations

import json
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import cs_port, ngm_e2e as e2e  # noqa: E402
from tests import cs_cases  # noqa: E402

CASES = [  # seed, read_len, n_reads, reference scale
    (21, 100, 6000, 1),
    (22, 150, 4200, 2),
    (23, 60, 2500, 1),
    (24, 250, 3000, 3),
]


def ngm_estimate(seed: int, read_len: int, n_reads: int, scale: int) -> float:
    contigs = cs_cases.make_reference(seed, scale)
    concat, ctg, _ = cs_port.layout(contigs)
    reads = cs_cases.make_reads(seed + 1, concat, ctg, n_reads, read_len, (read_len | 1) + 1)
    with tempfile.TemporaryDirectory(prefix="csens_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        e2e.run("ref", d, threads=2)
    return e2e.logged_sensitivity()


if __name__ == "__main__":
    out = [{"seed": s, "read_len": L, "n_reads": n, "scale": sc, "ngm_logged": "%f" % ngm_estimate(s, L, n, sc)} for s, L, n, sc in CASES]
    (Path(__file__).resolve().parent / "cs_sensitivity.json").write_text(json.dumps(out, indent=1) + "\n")
    print(out)
