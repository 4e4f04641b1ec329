"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = Path(__file__).resolve().parent / "golden"
FIELDS = ("pos", "qstart", "qend", "nm", "identity", "ascore", "cigar", "md")


def golden_names() -> List[str]:
    return sorted(p.stem for p in GOLDEN.glob("*.npz"))


def load_golden(name: str) -> Dict[str, np.ndarray]:
    with np.load(GOLDEN / f"{name}.npz", allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def scoring_kwargs(g) -> dict:
    m, x, gr, gf, bs, slam, tt, tc = (int(v) for v in g["scoring"])
    return dict(match=m, mismatch=x, gap_read=gr, gap_ref=gf, bs_mapping=bs, slam_seq=slam, match_tt=tt, match_tc=tc)


def bits(a) -> np.ndarray:
    """float32 -> raw bit pattern (NaN-safe exact comparison)."""
    return np.asarray(a, dtype=np.float32).view(np.uint32)


def align_tuple(pos, qstart, qend, nm, identity, ascore, cigar, md):
    return (int(pos), int(qstart), int(qend), int(nm), int(bits(identity)), float(ascore), bytes(cigar), bytes(md))


def golden_align_tuples(g, mode: int):
    n = len(g[f"pos{mode}"])
    return [align_tuple(g[f"pos{mode}"][i], g[f"qstart{mode}"][i], g[f"qend{mode}"][i], g[f"nm{mode}"][i],
                        g[f"identity{mode}"][i], g[f"ascore{mode}"][i], g[f"cigar{mode}"][i], g[f"md{mode}"][i])
            for i in range(n)]
