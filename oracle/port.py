"""oracle/port.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

ctypes front-end of ``oracle/libngm_oracle.so`` (plain-C restatement of the
reference's CPU alignment path, see ``oracle/ngm_oracle.h``).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libngm_oracle.so"


class _Params(C.Structure):
    _fields_ = [("match", C.c_float), ("mismatch", C.c_float), ("gap_read", C.c_float), ("gap_ref", C.c_float),
                ("match_alt", C.c_float), ("mismatch_alt", C.c_float), ("alt_scoring", C.c_int),
                ("bs_mapping", C.c_int), ("slam_seq", C.c_int), ("hard_clip", C.c_int), ("silent_clip", C.c_int)]


@dataclass
class Scoring:
    """Config.cpp:440-447 defaults; penalties positive like the config keys."""
    match: float = 10
    mismatch: float = 15
    gap_read: float = 20
    gap_ref: float = 20
    bs_mapping: int = 0
    slam_seq: int = 0
    match_tt: float = 0
    match_tc: float = 0
    hard_clip: int = 0
    silent_clip: int = 0

    def to_c(self) -> _Params:
        alt = 1 if self.bs_mapping == 1 else (2 if (self.slam_seq & 2) else 0)
        mm_alt = self.match_tc if alt == 1 else -self.match_tc
        return _Params(self.match, -self.mismatch, -self.gap_read, -self.gap_ref, self.match_tt, mm_alt, alt,
                       self.bs_mapping, self.slam_seq, self.hard_clip, self.silent_clip)


def build(force: bool = False) -> Path:
    if force or not LIB.exists() or LIB.stat().st_mtime < max(f.stat().st_mtime for f in HERE.glob("*_oracle.[ch]")):
        subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.ngm_oracle_batch_score.restype = C.c_int
        _lib.ngm_oracle_batch_align.restype = C.c_int
        _lib.ngm_oracle_decode_window.restype = C.c_int
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def batch_score(refs: np.ndarray, qrys: np.ndarray, qml: int, corridor: int, mode: int, sc: Scoring = Scoring(),
                dirs: Optional[np.ndarray] = None) -> np.ndarray:
    """refs uint8 [n, >=qml+corridor], qrys uint8 [n, >=qml] -> float32 [n] (SWOcl::BatchScore)."""
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    qrys = np.ascontiguousarray(qrys, dtype=np.uint8)
    n = refs.shape[0]
    assert refs.shape[1] >= qml + corridor and qrys.shape[1] >= qml and qrys.shape[0] == n
    out = np.full(n, np.nan, dtype=np.float32)
    p = sc.to_c()
    d = None if dirs is None else np.ascontiguousarray(dirs, dtype=np.uint8)
    got = lib().ngm_oracle_batch_score(C.byref(p), C.c_int(mode), C.c_int(qml), C.c_int(corridor), C.c_int(n),
                                       _ptr(refs), C.c_long(refs.shape[1]), _ptr(qrys), C.c_long(qrys.shape[1]),
                                       _ptr(d), _ptr(out))
    if got != n:
        out[:] = np.nan          # unsupported mode: nothing computed
    return out


@dataclass
class PortAlign:
    position_offset: int
    qstart: int
    qend: int
    nm: int
    identity: float
    ascore: float
    cigar: bytes      # as a C-string consumer sees it
    md: bytes         # as a C-string consumer sees it (truncated at an embedded NUL)
    md_raw: bytes     # every MD byte the reference wrote


def batch_align(refs: np.ndarray, qrys: np.ndarray, qml: int, corridor: int, mode: int, sc: Scoring = Scoring(),
                dirs: Optional[np.ndarray] = None) -> List[PortAlign]:
    """SWOclCigar::BatchAlign."""
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    qrys = np.ascontiguousarray(qrys, dtype=np.uint8)
    n = refs.shape[0]
    stride = 4 * max(1, qml) + 2 * corridor + 64
    pos = np.zeros(n, np.int32); qs = np.zeros(n, np.int32); qe = np.zeros(n, np.int32); nm = np.zeros(n, np.int32)
    ident = np.zeros(n, np.float32); asc = np.zeros(n, np.float32)
    cig = np.zeros((n, stride), np.uint8); md = np.zeros((n, stride), np.uint8)
    cig[:, :3] = 0x21; md[:, :3] = 0x21      # AlignmentBuffer.cpp:108-109 "!!!\0"
    cl = np.zeros(n, np.int32); ml = np.zeros(n, np.int32)
    p = sc.to_c()
    d = None if dirs is None else np.ascontiguousarray(dirs, dtype=np.uint8)
    lib().ngm_oracle_batch_align(C.byref(p), C.c_int(mode), C.c_int(qml), C.c_int(corridor), C.c_int(n),
                                 _ptr(refs), C.c_long(refs.shape[1]), _ptr(qrys), C.c_long(qrys.shape[1]), _ptr(d),
                                 _ptr(pos), _ptr(qs), _ptr(qe), _ptr(nm), _ptr(ident), _ptr(asc),
                                 _ptr(cig), _ptr(md), C.c_long(stride), _ptr(cl), _ptr(ml))
    out = []
    for i in range(n):
        cb = cig[i].tobytes(); mb = md[i].tobytes()
        raw = mb[: ml[i]] if ml[i] >= 0 else mb.split(b"\0")[0]
        out.append(PortAlign(int(pos[i]), int(qs[i]), int(qe[i]), int(nm[i]), float(ident[i]), float(asc[i]),
                             cb.split(b"\0")[0], mb.split(b"\0")[0], raw))
    return out


def pack_ref(ascii_ref: bytes) -> np.ndarray:
    n = len(ascii_ref)
    out = np.zeros((n + 1) // 2 + 2, dtype=np.uint8)
    lib().ngm_oracle_pack_ref(C.c_char_p(ascii_ref), C.c_ulonglong(n), _ptr(out))
    return out


def decode_window(packed: np.ndarray, concat_len: int, offset: int, buffer_len: int) -> Optional[bytes]:
    buf = np.full(buffer_len, 0x3F, dtype=np.uint8)
    ok = lib().ngm_oracle_decode_window(_ptr(packed), C.c_ulonglong(concat_len), C.c_ulonglong(offset),
                                        C.c_ulonglong(buffer_len), _ptr(buf))
    return buf.tobytes() if ok else None
