"""Seeded candidate-search test cases (test infrastructure): references with N runs, repeats and odd contig lengths;
reads with substitutions, indels, N's, both strands, some unmappable."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", np.uint8)
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def make_reference(seed: int, scale: int = 1) -> List[bytes]:
    rng = np.random.default_rng(seed)
    mk = lambda n: ACGT[rng.integers(0, 4, n)].tobytes()
    c1 = mk(60_001 * scale)
    c2 = mk(12_000 * scale) + b"N" * 37 + mk(13_000) + b"NNN" + mk(700) + b"N" + mk(9) + b"N" + mk(13) + b"NN"
    dup = mk(400)
    c3 = b"A" * 400 + mk(5_000 * scale) + b"ACAC" * 200 + mk(3_000) + dup + mk(1_500) + dup + mk(2_003)
    c4 = b"N" * 5 + mk(14) + b"N" + mk(2_000)
    return [c1, c2, c3, c4]


def make_reads(seed: int, concat: bytes, contigs: List[Tuple[int, int]], n_reads: int, read_len: int, stride: int) -> np.ndarray:
    """-> uint8 [n_reads, stride], NUL padded (MappedRead::Seq)."""
    rng = np.random.default_rng(seed)
    full = np.frombuffer(concat, np.uint8)
    out = np.zeros((n_reads, stride), np.uint8)
    for r in range(n_reads):
        L = read_len if r % 13 else int(rng.integers(12, read_len + 1))
        st, ln = contigs[r % len(contigs)]
        lo, hi = st - 20, st + ln - L + 20                       # some reads hang over the contig ends into the spacers
        pos = int(rng.integers(lo, max(lo + 1, hi)))
        s = full[pos: pos + L + 8].copy()
        kind = r % 17
        rate = 0.12 if kind == 4 else 0.03
        mut = rng.random(len(s)) < rate
        s[mut] = ACGT[rng.integers(0, 4, int(mut.sum()))]
        if kind == 3:
            s[int(rng.integers(0, L))] = ord("N")
        if kind == 5 and L > 50:
            s = np.concatenate([s[:40], s[43:]])
        if kind == 6 and L > 50:
            s = np.concatenate([s[:31], ACGT[rng.integers(0, 4, 2)], s[31:]])
        if kind == 7:
            s = ACGT[rng.integers(0, 4, L)]                      # unmappable
        if kind == 8:
            s[:] = ord("N")
        if kind == 9:
            s[: L // 2] = ord("N")
        s = s[:L]
        b = s.tobytes()
        if r & 1:
            b = b.translate(COMP)[::-1]
        out[r, : len(b)] = np.frombuffer(b, np.uint8)
    return out


def convert_bases(reads: np.ndarray, seed: int, mode: int, paired: bool = False) -> np.ndarray:
    """Chemistry of the mutated searches: mode 1 (bisulfite) turns ~80 % of the C into T -- second mates: G into A --, mode 2 (SLAMseq) ~5 % of the
    T into C -- second mates: A into G; the k-mer mutation of CS::PrefixMutateSearch undoes exactly these."""
    rng = np.random.default_rng(seed)
    out = reads.copy()
    second = np.zeros(out.shape, bool)
    if paired:
        second[1::2] = True
    draw = rng.random(out.shape)
    if mode == 1:
        out[(reads == ord("C")) & ~second & (draw < 0.8)] = ord("T")
        out[(reads == ord("G")) & second & (draw < 0.8)] = ord("A")
    else:
        out[(reads == ord("T")) & ~second & (draw < 0.05)] = ord("C")
        out[(reads == ord("A")) & second & (draw < 0.05)] = ord("G")
    return out


def write_fasta(path, contig_seqs: List[bytes]) -> None:
    with open(path, "wb") as f:
        for i, s in enumerate(contig_seqs):
            f.write(b">c%d\n" % i)
            for j in range(0, len(s), 70):
                f.write(s[j: j + 70] + b"\n")


def write_fastq(path, reads: np.ndarray) -> None:
    with open(path, "wb") as f:
        for r in range(reads.shape[0]):
            s = reads[r].tobytes().split(b"\0")[0]
            f.write(b"@r%d\n" % r + s + b"\n+\n" + b"I" * len(s) + b"\n")
