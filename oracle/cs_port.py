"""oracle/cs_port.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

ctypes front-end of the candidate-search restatement in ``oracle/cs_oracle.c`` plus drivers for the reference's own
code: ``run_probe`` (oracle/_ref/ngm/ngm_cs_probe: the reference's CS on a read file) and ``read_ht_file``
(``<ref>-ht-<k>-<skip>.3.ngm`` as written by CompactPrefixTable::saveToFile, PrefixTable.cpp:819-859).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

from oracle import port

HERE = Path(__file__).resolve().parent
PROBE = HERE / "_ref" / "ngm" / "ngm_cs_probe"


class _Contig(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_uint64)]


class _Index(C.Structure):
    _fields_ = [("k", C.c_int), ("ref_skip", C.c_int), ("bin_shift", C.c_int), ("index_len", C.c_uint32), ("table_len", C.c_uint32),
                ("tab", C.POINTER(C.c_uint32)), ("weight", C.POINTER(C.c_int8)), ("table", C.POINTER(C.c_uint32)), ("max_kfreq", C.c_int)]


CAND = np.dtype([("location", "<u8"), ("score", "<f4"), ("reverse", "<i4")])


def _lib():
    lib = port.lib()
    lib.cs_oracle_build_index.restype = C.c_int
    lib.cs_oracle_search_batch.restype = C.c_longlong
    lib.cs_oracle_revcomp.restype = C.c_uint32
    return lib


class Index:
    """CompactPrefixTable (one unit) built by the restatement."""

    def __init__(self, packed: np.ndarray, concat_len: int, contigs: Sequence[Tuple[int, int]], k: int = 13, ref_skip: int = 2, bin_shift: int = 2,
                 skip_rep: bool = True):
        self.lib = _lib()
        self.c = _Index()
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        arr = (_Contig * len(contigs))(*[_Contig(s, l) for s, l in contigs])
        rc = self.lib.cs_oracle_build_index(packed.ctypes.data_as(C.c_void_p), C.c_uint64(concat_len), arr, len(contigs), k, ref_skip, bin_shift,
                                            1 if skip_rep else 0, C.byref(self.c))
        if rc != 0:
            raise ValueError("cs_oracle_build_index failed")
        self.k, self.ref_skip, self.bin_shift = k, ref_skip, bin_shift
        self.index_len, self.table_len, self.max_kfreq = int(self.c.index_len), int(self.c.table_len), int(self.c.max_kfreq)
        self.tab = np.ctypeslib.as_array(self.c.tab, shape=(self.index_len,)).copy()
        self.weight = np.ctypeslib.as_array(self.c.weight, shape=(self.index_len,)).copy()
        self.table = np.ctypeslib.as_array(self.c.table, shape=(max(self.table_len, 1),)).copy()[: self.table_len]

    @classmethod
    def from_arrays(cls, tab: np.ndarray, weight: np.ndarray, table: np.ndarray, k: int, ref_skip: int = 2, bin_shift: int = 2, max_kfreq: int = 100):
        """Wrap an existing prefix table (e.g. read from an ht file or exported from the device) for `search`."""
        self = cls.__new__(cls)
        self.lib = _lib()
        self.tab = np.ascontiguousarray(tab, dtype=np.uint32)
        self.weight = np.ascontiguousarray(weight, dtype=np.int8)
        self.table = np.ascontiguousarray(table, dtype=np.uint32)
        self.k, self.ref_skip, self.bin_shift = k, ref_skip, bin_shift
        self.index_len, self.table_len, self.max_kfreq = len(self.tab), len(self.table), max_kfreq
        self.c = _Index(k, ref_skip, bin_shift, self.index_len, self.table_len, self.tab.ctypes.data_as(C.POINTER(C.c_uint32)),
                        self.weight.ctypes.data_as(C.POINTER(C.c_int8)), self.table.ctypes.data_as(C.POINTER(C.c_uint32)), max_kfreq)
        self._borrowed = True
        return self

    def search(self, reads: np.ndarray, sensitivity: float, kmer_min: float = 0.0, max_kfreq: int = 0, max_cmrs: int = 2 ** 31 - 1):
        """reads uint8 [n, stride] NUL padded -> (cand_begin int32 [n+1], candidates CAND [total], max_hit float32 [n])."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, stride = reads.shape
        begin = np.zeros(n + 1, np.int32)
        mh = np.zeros(n, np.float32)
        cap = max(1024, 64 * n)
        while True:
            out = np.zeros(cap, dtype=CAND)
            total = self.lib.cs_oracle_search_batch(C.byref(self.c), reads.ctypes.data_as(C.c_void_p), n, stride, C.c_float(sensitivity), C.c_float(kmer_min),
                                                    max_kfreq or self.max_kfreq, max_cmrs, begin.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                                    C.c_longlong(cap), mh.ctypes.data_as(C.c_void_p))
            if total <= cap:
                return begin, out[:total], mh
            cap = int(total) + 16

    def search_mut(self, reads: np.ndarray, sensitivity: float, mutate_mode: int, bs_cutoff: int = 6, paired: bool = False, read_skip: int = 0,
                   table_bits: int = 16, kmer_min: float = 0.0, max_kfreq: int = 0, max_cmrs: int = 2 ** 31 - 1):
        """CS::RunBatch under --bs-mapping (mutate_mode 1) / --slam-seq 4 (mutate_mode 2); same outputs as `search`."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, stride = reads.shape
        begin = np.zeros(n + 1, np.int32)
        mh = np.zeros(n, np.float32)
        cap = max(1024, 64 * n)
        self.lib.cs_oracle_search_batch_mut.restype = C.c_longlong
        while True:
            out = np.zeros(cap, dtype=CAND)
            total = self.lib.cs_oracle_search_batch_mut(C.byref(self.c), reads.ctypes.data_as(C.c_void_p), n, stride, C.c_float(sensitivity),
                                                        C.c_float(kmer_min), max_kfreq or self.max_kfreq, max_cmrs, mutate_mode, bs_cutoff,
                                                        1 if paired else 0, read_skip, table_bits, begin.ctypes.data_as(C.c_void_p),
                                                        out.ctypes.data_as(C.c_void_p), C.c_longlong(cap), mh.ctypes.data_as(C.c_void_p))
            if total <= cap:
                return begin, out[:total], mh
            cap = int(total) + 16

    def estimate_sensitivity(self, reads: np.ndarray, max_kfreq: int = 0):
        """ReadProvider::init's estimate over the whole input (every 1000th read of the first 10 M) -> (sensitivity, contributing reads);
        (0.5, 0) for fewer than 1000 reads (ReadProvider.cpp:310,372-379)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n = reads.shape[0]
        if n < 1000:
            return 0.5, 0
        sample = np.ascontiguousarray(reads[999:min(n, 10_000_000 - 1):1000])
        out = C.c_float(0.0)
        self.lib.cs_oracle_estimate_sensitivity.restype = C.c_int
        cnt = self.lib.cs_oracle_estimate_sensitivity(C.byref(self.c), sample.ctypes.data_as(C.c_void_p), sample.shape[0], sample.shape[1],
                                                      max_kfreq or self.max_kfreq, C.byref(out))
        return float(out.value), int(cnt)

    def close(self):
        if getattr(self, "_borrowed", False):
            return
        if self.c.tab:
            self.lib.cs_oracle_free_index(C.byref(self.c))


def revcomp_prefix(prefix: int, k: int) -> int:
    return int(_lib().cs_oracle_revcomp(C.c_uint32(prefix), k))


def layout(contig_seqs: Sequence[bytes]):
    """SequenceProvider::Init layout (SequenceProvider.cpp:289-330): 1000 N, contig (+ 1 N if odd), 1000 N, ...
    -> (concatenated ASCII, [(start, length)], concat_len = len - 1)."""
    concat = b"N" * 1000
    contigs = []
    for s in contig_seqs:
        contigs.append((len(concat), len(s)))
        concat += s + (b"N" if len(s) & 1 else b"") + b"N" * 1000
    return concat, contigs, len(concat) - 1


def ngm_env():
    env = dict(os.environ)
    ocl = HERE / "_ref" / "ocl"
    env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
    env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def probe_available() -> bool:
    return PROBE.exists()


def run_probe(d: Path, ref: str, reads: str, sensitivity: float, k: int = 13, extra: Sequence[str] = ()):
    """-> (header dict, list of (read_id, name, length, max_hit, [(location, reverse, votes)]))"""
    cmd = [str(PROBE), "-r", ref, "-q", reads, "-s", repr(float(sensitivity)), "-k", str(k), "-t", "1", "--no-progress", *extra]
    p = subprocess.run(cmd, env=ngm_env(), capture_output=True, text=True, cwd=d)
    if p.returncode != 0:
        raise RuntimeError(f"ngm_cs_probe failed ({p.returncode}):\n{p.stdout[-800:]}\n{p.stderr[-1500:]}")
    head, rows = {}, []
    for ln in p.stdout.splitlines():
        f = ln.split()
        if ln.startswith("#"):
            head = {"max_kfreq": int(f[1]), "sensitivity": float(f[3]), "kmer": int(f[5])}
            continue
        n = int(f[4])
        cands = []
        for tok in f[5: 5 + n]:
            loc, rev, votes = tok.split(":")
            cands.append((int(loc), int(rev), float(votes)))
        rows.append((int(f[0]), f[1], int(f[2]), float(f[3]), cands))
    return head, rows


def read_ht_file(path) -> dict:
    """CompactPrefixTable::saveToFile layout (PrefixTable.cpp:819-859); Index is a packed {uint32, char} (PrefixTable.h:20-33)."""
    raw = np.fromfile(path, dtype=np.uint8)
    hdr = raw[:20].view("<u4")
    cookie, k, skip, units, index_len = (int(x) for x in hdr)
    assert cookie == 0x74656 and units == 1
    off = 20
    table_len = int(raw[off: off + 4].view("<u4")[0])
    off += 4
    idx = raw[off: off + 5 * index_len].view(np.dtype([("tab", "<u4"), ("weight", "i1")]))
    off += 5 * index_len
    table = raw[off: off + 4 * table_len].view("<u4")
    off += 4 * table_len
    unit_offset = int(raw[off: off + 8].view("<u8")[0])
    off += 8
    sig = int(raw[off: off + 4].view("<u4")[0])
    assert sig == (cookie + k + skip + units + index_len) & 0xFFFFFFFF
    return {"k": k, "ref_skip": skip, "index_len": index_len, "table_len": table_len, "tab": idx["tab"].copy(), "weight": idx["weight"].copy(),
            "table": table.copy(), "offset": unit_offset}
