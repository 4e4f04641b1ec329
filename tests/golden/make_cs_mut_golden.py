"""Generate tests/golden/cs_mut/*.npz by running the UNMODIFIED reference's candidate search under --bs-mapping and --slam-seq.

    python __graft_entry__.py            # builds oracle/_ref/ngm/ngm_cs_probe
    python tests/golden/make_cs_mut_golden.py

oracle/_ref/ngm/ngm_cs_probe drives the reference's own CS::PrefixIteration / PrefixMutateSearch / PrefixSearch / AddLocationStd /
CollectResultsStd with CS::RunBatch's choice of mutated base per mate and its table-overflow retries.  A fixture holds the seeded
inputs and the per-read candidate lists in the reference's order (votes are fractional under --slam-seq).
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

OUT = Path(__file__).resolve().parent / "cs_mut"


def emit(name: str, seed: int, k: int, read_len: int, n_reads: int, sensitivity: float, mode: int, paired: bool, bs_cutoff: int = 6):
    contigs = cs_cases.make_reference(seed)
    concat, ctg, concat_len = cs_port.layout(contigs)
    stride = (read_len | 1) + 1
    reads = cs_cases.convert_bases(cs_cases.make_reads(seed + 1, concat, ctg, n_reads, read_len, stride), seed + 2, mode, paired)
    extra = ["--bs-mapping", "--bs-cutoff", str(bs_cutoff)] if mode == 1 else ["--slam-seq", "4"]
    if paired:
        extra += ["-p", "--skip-mate-check"]
    ref_skip = 0 if mode == 1 else 2                                 # CompactPrefixTable under bs_mapping, PrefixTable.cpp:204-207
    with tempfile.TemporaryDirectory(prefix="csmutgold_") as td:
        d = Path(td)
        cs_cases.write_fasta(d / "ref.fa", contigs)
        cs_cases.write_fastq(d / "reads.fq", reads)
        head, rows = cs_port.run_probe(d, "ref.fa", "reads.fq", sensitivity, k=k, extra=extra)
        assert (d / f"ref.fa-ht-{k}-{ref_skip}.3.ngm").exists()
    idx = np.array([int(r[1][1:]) for r in rows], np.int32)
    begin = np.zeros(len(rows) + 1, np.int32)
    loc, rev, votes = [], [], []
    for i, r in enumerate(rows):
        for (l, s, v) in r[4]:
            loc.append(l)
            rev.append(s)
            votes.append(v)
        begin[i + 1] = len(loc)
    np.savez_compressed(OUT / f"{name}.npz", seed=seed, k=k, read_len=read_len, sensitivity=np.float32(sensitivity), max_kfreq=head["max_kfreq"],
                        mode=mode, paired=int(paired), bs_cutoff=bs_cutoff, ref_skip=ref_skip, read_skip=2 if mode == 1 else 0,
                        concat=np.frombuffer(concat, np.uint8), contigs=np.array(ctg, np.int64), reads=reads, read_index=idx,
                        max_hit=np.array([r[3] for r in rows], np.float32), cand_begin=begin, cand_loc=np.array(loc, np.uint64),
                        cand_rev=np.array(rev, np.uint8), cand_votes=np.array(votes, np.float32))
    frac = int(np.sum(np.array(votes) != np.floor(votes)))
    print(name, "reads", len(rows), "candidates", len(loc), "fractional votes", frac, "max_kfreq", head["max_kfreq"])


if __name__ == "__main__":
    OUT.mkdir(exist_ok=True)
    emit("bs_k10_l75", 111, 10, 75, 400, 0.5, 1, False)
    emit("bs_k12_l100_pe_cut3", 222, 12, 100, 400, 0.6, 1, True, bs_cutoff=3)
    emit("slam_k12_l100", 333, 12, 100, 400, 0.5, 2, False)
    emit("slam_k13_l150_pe", 444, 13, 150, 400, 0.3, 2, True)
