"""oracle/ref_driver.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Runs the unmodified reference backend (``oracle/_ref/ngm_ref_harness``, built by
``oracle/Makefile`` from /root/reference/lib/mason/opencl/*.cpp) on a batch of
(ref window, read) pairs and returns what ``SWOclCigar::BatchScore`` /
``BatchAlign`` produced.  The harness executes the reference's CPU OpenCL
kernels on the AMD APP runtime vendored by the reference (copied to
``oracle/_ref/ocl``), so nothing here needs /root/reference at run time.
"""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
HARNESS = REF_DIR / "ngm_ref_harness"
MAGIC = 0x4E474D31


@dataclass
class Scoring:
    """Linear-gap scoring, Config.cpp:440-447 defaults (penalties are positive)."""
    match: int = 10
    mismatch: int = 15
    gap_read: int = 20
    gap_ref: int = 20
    bs_mapping: int = 0
    slam_seq: int = 0
    match_tt: int = 0
    match_tc: int = 0


@dataclass
class RefAlign:
    position_offset: int
    qstart: int
    qend: int
    nm: int
    identity: float
    ascore: float
    cigar: bytes
    md: bytes


@dataclass
class RefResult:
    scores: np.ndarray
    aligns: List[RefAlign] = field(default_factory=list)


def available() -> bool:
    return HARNESS.exists() and (REF_DIR / "ocl" / "lib" / "libamdocl64.so").exists()


def _env() -> dict:
    env = dict(os.environ)
    env["OPENCL_VENDOR_PATH"] = str(REF_DIR / "ocl" / "vendor")
    # LD_LIBRARY_PATH is mandatory for the AMD ICD to find libamdocl64.so
    env["LD_LIBRARY_PATH"] = str(REF_DIR / "ocl" / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def ref_buf_len_score(qml: int, corridor: int) -> int:
    """ScoreBuffer.h:112 -- refMaxLen = ((qml + corridor) | 1) + 1."""
    return ((qml + corridor) | 1) + 1


def pack_fixed(seqs: Sequence[bytes], width: int) -> np.ndarray:
    """NUL-padded fixed-width rows, like MappedRead::Seq / the decoded window buffers."""
    out = np.zeros((len(seqs), width), dtype=np.uint8)
    for i, s in enumerate(seqs):
        if len(s) > width:
            raise ValueError(f"sequence {i} longer ({len(s)}) than buffer ({width})")
        out[i, : len(s)] = np.frombuffer(s, dtype=np.uint8)
    return out


def write_job(path: Path, refs: np.ndarray, qrys: np.ndarray, qml: int, corridor: int, mode: int,
              sc: Scoring, n_align: int, dirs: Optional[np.ndarray]) -> None:
    n, rbl = refs.shape
    assert qrys.shape == (n, qml)
    hdr = [MAGIC, qml, corridor, mode, n, sc.match, sc.mismatch, sc.gap_read, sc.gap_ref, sc.bs_mapping,
           sc.slam_seq, sc.match_tt, sc.match_tc, rbl, n_align, 1 if dirs is not None else 0]
    with open(path, "wb") as f:
        f.write(struct.pack("<16i", *hdr))
        f.write(np.ascontiguousarray(refs, dtype=np.uint8).tobytes())
        f.write(np.ascontiguousarray(qrys, dtype=np.uint8).tobytes())
        if dirs is not None:
            f.write(np.ascontiguousarray(dirs, dtype=np.uint8).tobytes())


def run(refs: np.ndarray, qrys: np.ndarray, qml: int, corridor: int, mode: int,
        sc: Scoring = Scoring(), align: bool = True, dirs: Optional[np.ndarray] = None) -> RefResult:
    """refs: uint8 [n, ref_buf_len] (>= qml+corridor), qrys: uint8 [n, qml]."""
    if not available():
        raise RuntimeError("oracle/_ref not built: run `make -C oracle ref` where /root/reference exists")
    n = refs.shape[0]
    with tempfile.TemporaryDirectory(prefix="ngmref_") as td:
        inp, out = Path(td) / "in.bin", Path(td) / "out.bin"
        write_job(inp, refs, qrys, qml, corridor, mode, sc, n if align else 0, dirs)
        p = subprocess.run([str(HARNESS), "run", str(inp), str(out)], env=_env(), capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"reference harness failed ({p.returncode}): {p.stderr[-2000:]}")
        raw = out.read_bytes()
    scores = np.frombuffer(raw, dtype="<f4", count=n).copy()
    res = RefResult(scores=scores)
    off = 4 * n
    if align:
        for _ in range(n):
            pos, qs, qe, nm, ident, asc, clen, mlen = struct.unpack_from("<4i2f2H", raw, off)
            off += 28
            cigar = raw[off: off + clen]
            off += clen
            md = raw[off: off + mlen]
            off += mlen
            res.aligns.append(RefAlign(pos, qs, qe, nm, ident, asc, cigar, md))
    return res


def bench(refs: np.ndarray, qrys: np.ndarray, qml: int, corridor: int, mode: int, n_align: int,
          threads: int, warm_passes: int, timed_passes: int, sc: Scoring = Scoring()) -> dict:
    """Time BatchScore over all pairs + BatchAlign over the first n_align pairs with `threads` CS-thread
    equivalents (one SWOclCigar instance on its own 1-core OpenCL sub-device each, like `ngm -t`).
    Instance construction (OpenCL JIT) and the warm passes are outside the reported wall time."""
    if not available():
        raise RuntimeError("oracle/_ref not built")
    with tempfile.TemporaryDirectory(prefix="ngmref_") as td:
        inp = Path(td) / "in.bin"
        write_job(inp, refs, qrys, qml, corridor, mode, sc, n_align, None)
        p = subprocess.run([str(HARNESS), "bench", str(inp), str(threads), str(warm_passes), str(timed_passes)], env=_env(),
                           capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"reference harness failed ({p.returncode}): {p.stderr[-2000:]}")
    return json.loads(p.stdout.strip().splitlines()[-1])
