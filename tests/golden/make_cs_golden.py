"""Generate tests/golden/cs/*.npz by running the UNMODIFIED reference's candidate search.

    python __graft_entry__.py            # builds oracle/_ref/ngm/{ngm_ref,ngm_cs_probe}
    python tests/golden/make_cs_golden.py

Every fixture holds seeded inputs (contigs, reads) and what the reference produced for them: the prefix-table file
``<ref>-ht-<k>-2.3.ngm`` written by CompactPrefixTable::saveToFile (index, weights, position table), max_kfreq, and the
per-read candidate lists of CS::PrefixSearch/AddLocationStd/CollectResultsStd in their original order
(oracle/_ref/ngm/ngm_cs_probe).  NGM skips reads shorter than ... none here: the probe sees every read.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import cs_port  # noqa: E402
from tests import cs_cases  # noqa: E402

OUT = Path(__file__).resolve().parent / "cs"


def emit(name: str, seed: int, k: int, read_len: int, n_reads: int, sensitivity: float):
    contigs = cs_cases.make_reference(seed)
    concat, ctg, concat_len = cs_port.layout(contigs)
    stride = (read_len | 1) + 1
    reads = cs_cases.make_reads(seed + 1, concat, ctg, n_reads, read_len, stride)
    with tempfile.TemporaryDirectory(prefix="csgold_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        head, rows = cs_port.run_probe(d, "ref.fa", "reads.fq", sensitivity, k=k)
        ht = cs_port.read_ht_file(d / f"ref.fa-ht-{k}-2.3.ngm")
    # NGM's read parser drops nothing here, but keep the mapping explicit: probe row -> read index by name
    idx = np.array([int(r[1][1:]) for r in rows], np.int32)
    begin = np.zeros(len(rows) + 1, np.int32)
    loc, rev, votes = [], [], []
    for i, r in enumerate(rows):
        for (l, s, v) in r[4]:
            loc.append(l)
            rev.append(s)
            votes.append(v)
        begin[i + 1] = len(loc)
    used = ht["weight"] != 0
    np.savez_compressed(OUT / f"{name}.npz", seed=seed, k=k, read_len=read_len, sensitivity=np.float32(sensitivity), max_kfreq=head["max_kfreq"],
                        concat=np.frombuffer(concat, np.uint8), contigs=np.array(ctg, np.int64), reads=reads, read_index=idx,
                        read_length=np.array([r[2] for r in rows], np.int32), max_hit=np.array([r[3] for r in rows], np.float32), cand_begin=begin,
                        cand_loc=np.array(loc, np.uint64), cand_rev=np.array(rev, np.uint8), cand_votes=np.array(votes, np.float32),
                        ht_table_len=ht["table_len"], ht_used_prefix=np.nonzero(used)[0].astype(np.uint32), ht_used_weight=ht["weight"][used],
                        ht_used_count=(ht["tab"][1:][used[:-1]] - ht["tab"][:-1][used[:-1]]).astype(np.uint32), ht_table=ht["table"])
    print(name, "reads", len(rows), "candidates", len(loc), "table_len", ht["table_len"], "max_kfreq", head["max_kfreq"])


if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    emit("k10_l100", 101, 10, 100, 1200, 0.5)
    emit("k12_l150", 202, 12, 150, 800, 0.3)
    emit("k13_l150", 303, 13, 150, 800, 0.5)
    emit("k11_l250", 404, 11, 250, 500, 0.7)
