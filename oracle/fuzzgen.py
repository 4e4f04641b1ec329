"""oracle/fuzzgen.py -- TEST INFRASTRUCTURE ONLY.

Seeded generator of (reference window, read) pairs that exercise the corners
of the reference's banded DP (SURVEY 8a parity notes): substitutions, indels
up to the band edge, read/ref N, 'x' and NUL window tails, short and empty
reads, unrelated reads, lower-case bases.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def ref_buf_len_score(qml: int, corridor: int) -> int:
    """ScoreBuffer.h:112"""
    return ((qml + corridor) | 1) + 1


def mutate(seq: np.ndarray, rng: np.random.Generator, sub: float, ins: float, dele: float) -> np.ndarray:
    out = []
    for b in seq:
        r = rng.random()
        if r < dele:
            continue
        if r < dele + ins:
            k = int(rng.integers(1, 4))
            out.extend(ACGT[rng.integers(0, 4, k)])
        if rng.random() < sub:
            out.append(ACGT[rng.integers(0, 4)])
        else:
            out.append(b)
    return np.array(out, dtype=np.uint8)


def make_pairs(n: int, qml: int, corridor: int, seed: int, clean: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (refs uint8 [n, ref_buf_len], qrys uint8 [n, qml]).

    clean=True restricts to ACGT-only windows and reads of near-maximal length (the
    common case on the hot path); otherwise every 4th pair or so gets an edge case.
    """
    rng = np.random.default_rng(seed)
    rbl = ref_buf_len_score(qml, corridor)
    refs = np.zeros((n, rbl), dtype=np.uint8)
    qrys = np.zeros((n, qml), dtype=np.uint8)
    max_read = qml - 1 if qml > 1 else 0
    for i in range(n):
        wlen = qml + corridor
        win = ACGT[rng.integers(0, 4, wlen)]
        kind = 0 if clean else int(rng.integers(0, 12))
        # read length: mostly near max (qml = (L|1)+1 leaves 1-2 pad rows), sometimes short
        if kind == 1:
            L = int(rng.integers(1, max(2, max_read + 1)))
        elif kind == 2 and not clean:
            L = int(rng.integers(0, 3))                    # empty / 1-2 bp reads
        else:
            L = max(1, max_read - int(rng.integers(0, 2)))
        start = corridor // 2 + int(rng.integers(-min(3, corridor // 2), min(3, corridor // 2) + 1))
        start = max(0, start)
        src = win[start:start + L + 8]
        err = float(rng.choice([0.0, 0.01, 0.03, 0.1, 0.3]))
        indel = float(rng.choice([0.0, 0.0, 0.002, 0.01, 0.03]))
        read = mutate(src, rng, err, indel, indel)[:L]
        if kind == 3:
            read = ACGT[rng.integers(0, 4, L)]              # unrelated read
        if kind == 4 and L > 20:                            # one long indel near the band limit
            cut = int(rng.integers(5, L - 5))
            glen = int(rng.integers(1, max(2, corridor // 2 + 2)))
            if rng.random() < 0.5:
                read = np.concatenate([read[:cut], read[cut + glen:]])
            else:
                read = np.concatenate([read[:cut], ACGT[rng.integers(0, 4, glen)], read[cut:]])[:L]
        if kind == 5 and len(read):                         # N in read
            for _ in range(int(rng.integers(1, 4))):
                read[int(rng.integers(0, len(read)))] = ord("N")
        if kind == 6:                                       # N in window
            for _ in range(int(rng.integers(1, 6))):
                win[int(rng.integers(0, wlen))] = ord("N")
        if kind == 7:                                       # 'x' tail (beyond the concatenated reference)
            k = int(rng.integers(1, corridor + 12))
            win[wlen - k:] = ord("x")
        if kind == 8:                                       # NUL tail (window shorter than the buffer)
            k = int(rng.integers(1, corridor + 12))
            win[wlen - k:] = 0
        if kind == 9 and len(read):                         # lower case, stray characters
            j = int(rng.integers(0, len(read)))
            read[j] = ord(chr(read[j]).lower())
            win[int(rng.integers(0, wlen))] = ord("a")
            if rng.random() < 0.3:
                read[int(rng.integers(0, len(read)))] = ord("R")
        if kind == 10:                                      # N vs N
            j = int(rng.integers(0, max(1, len(read))))
            if len(read):
                read[j] = ord("N")
                win[min(wlen - 1, start + j)] = ord("N")
        if kind == 11:                                      # low-complexity
            read[:] = ord("A")
            if rng.random() < 0.5:
                win[:] = ACGT[rng.integers(0, 2, wlen)]
        refs[i, :wlen] = win
        # the byte at index qml+corridor..rbl-1 stays NUL like DecodeRefSequence's tail
        qrys[i, :len(read)] = read
    return refs, qrys
