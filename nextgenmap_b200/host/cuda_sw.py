"""Host-side mirror of the reference's alignment plugin interface on top of the C ABI.

``CudaSW`` has the surface of the reference's ``IAlignment`` (include/IAlignment.h:48-69)
as implemented by ``SWOclCigar`` (lib/mason/opencl/SWOclCigar.h:17-42): ``GetScoreBatchSize``,
``GetAlignBatchSize``, ``BatchScore``, ``BatchAlign`` -- same argument meaning, same result
fields (``Align``: pBuffer1 = CIGAR, pBuffer2 = MD, PositionOffset, QStart, QEnd, Score,
Identity, NM) -- plus the descriptor fast path.  Everything is computed by
``libngm_b200.so`` (hand-written sm_100a CUDA); there is no CPU fallback: if the library or a
CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

PKG = Path(__file__).resolve().parents[1]
LIB_PATH = Path(os.environ["NGM_B200_LIB"]) if os.environ.get("NGM_B200_LIB") else PKG / "libngm_b200.so"      # (A/B builds: see build.py)


class NgmB200Error(RuntimeError):
    pass


class Params(C.Structure):
    """ngm_b200_params (include/ngm_b200.h); defaults = Config.cpp:440-447."""
    _fields_ = [("qry_max_len", C.c_int32), ("corridor", C.c_int32), ("match_bonus", C.c_float),
                ("mismatch_penalty", C.c_float), ("gap_read_penalty", C.c_float), ("gap_ref_penalty", C.c_float),
                ("match_bonus_tt", C.c_float), ("match_bonus_tc", C.c_float), ("bs_mapping", C.c_int32),
                ("slam_seq", C.c_int32), ("hard_clip", C.c_int32), ("silent_clip", C.c_int32), ("device", C.c_int32),
                ("score_batch", C.c_int32), ("align_batch", C.c_int32), ("lane_mode", C.c_int32)]


class CsParams(C.Structure):
    """ngm_b200_cs_params (include/ngm_b200.h); defaults = src/config/Config.cpp:383-390,512."""
    _fields_ = [("kmer", C.c_int32), ("kmer_skip", C.c_int32), ("bin_size", C.c_int32), ("skip_rep", C.c_int32), ("sensitivity", C.c_float),
                ("kmer_min", C.c_float), ("max_kfreq", C.c_int32), ("max_cmrs", C.c_int32)]


class PeParams(C.Structure):
    """ngm_b200_pe_params (include/ngm_b200.h); defaults = src/config/Config.cpp:393-406."""
    _fields_ = [("pair_score_cutoff", C.c_float), ("min_insert_size", C.c_int32), ("max_insert_size", C.c_int32), ("strata", C.c_int32),
                ("fast_pairing", C.c_int32)]


class SamOpts(C.Structure):
    """ngm_b200_sam_opts (include/ngm_b200.h); defaults = src/config/Config.cpp:405-410,430-431."""
    _fields_ = [("min_identity", C.c_float), ("min_residues", C.c_float), ("min_insert_size", C.c_int32), ("max_insert_size", C.c_int32), ("threads", C.c_int32),
                ("min_mq", C.c_int32), ("clip_seq", C.c_int32), ("read_group", C.c_char_p), ("bs_mapping", C.c_int32), ("slam_seq", C.c_int32)]


class SamBatch(C.Structure):
    """ngm_b200_sam_batch (include/ngm_b200.h)."""
    _fields_ = [("n_reads", C.c_int32), ("stride", C.c_int32), ("reads", C.c_void_p), ("quals", C.c_void_p), ("names", C.POINTER(C.c_char_p)), ("pairs", C.c_void_p),
                ("scores", C.c_void_p), ("best_pair", C.c_void_p), ("mapq", C.c_void_p), ("num_top", C.c_void_p), ("pair_fail", C.c_void_p), ("max_hit", C.c_void_p),
                ("recs", C.c_void_p), ("strings", C.c_void_p), ("topn", C.c_int32), ("sel", C.c_void_p), ("n_sel", C.c_void_p)]


class MapResult(C.Structure):
    """ngm_b200_map_result (include/ngm_b200.h)."""
    _fields_ = [("cand_begin", C.c_void_p), ("pairs", C.c_void_p), ("scores", C.c_void_p), ("capacity", C.c_size_t), ("n_candidates", C.c_size_t),
                ("best_pair", C.c_void_p), ("mapq", C.c_void_p), ("num_top", C.c_void_p), ("pair_fail", C.c_void_p), ("max_hit", C.c_void_p),
                ("recs", C.c_void_p), ("strings", C.c_void_p), ("str_capacity", C.c_size_t), ("str_used", C.c_size_t), ("sel", C.c_void_p), ("n_sel", C.c_void_p)]


class BatchIn(C.Structure):
    """ngm_b200_batch_in (include/ngm_b200.h)."""
    _fields_ = [("n_reads", C.c_int32), ("mode", C.c_int32), ("paired", C.c_int32), ("read_format", C.c_int32), ("reads", C.c_void_p), ("read_stride", C.c_int32),
                ("desc_format", C.c_int32), ("read_len", C.c_void_p), ("exceptions", C.c_void_p), ("n_exceptions", C.c_uint32), ("n_desc", C.c_uint32),
                ("cand_begin", C.c_void_p), ("desc", C.c_void_p)]


class BatchOut(C.Structure):
    """ngm_b200_batch_out (include/ngm_b200.h)."""
    _fields_ = [("scores", C.c_void_p), ("best_pair", C.c_void_p), ("mapq", C.c_void_p), ("num_top", C.c_void_p), ("pair_fail", C.c_void_p), ("recs", C.c_void_p),
                ("strings", C.c_void_p), ("str_capacity", C.c_size_t), ("str_used", C.c_size_t), ("d_str_cursor", C.c_void_p), ("sel", C.c_void_p),
                ("n_sel", C.c_void_p)]


READ_EXC = np.dtype([("read_index", "<u4"), ("pos", "<u2"), ("ch", "u1"), ("pad", "u1")])
READS_ASCII, READS_PACKED2 = 0, 1
DESC_PAIR16, DESC_U64 = 0, 1


def make_desc_u64(window_start, flags) -> np.ndarray:
    """NGM_B200_DESC(): window start saturated to 56 bits, flags in the top byte."""
    ws = np.minimum(np.asarray(window_start, dtype=np.uint64), np.uint64(0x00FFFFFFFFFFFFFF))
    return ws | (np.asarray(flags, dtype=np.uint64) << np.uint64(56))


class _CContigRec(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_uint32), ("name_len", C.c_uint32), ("name", C.c_char * 100)]


class _CHtFile(C.Structure):
    _fields_ = [("kmer", C.c_uint32), ("kmer_skip", C.c_uint32), ("index_len", C.c_uint32), ("table_len", C.c_uint32), ("tab", C.POINTER(C.c_uint32)),
                ("weight", C.POINTER(C.c_int8)), ("table", C.POINTER(C.c_uint32)), ("unit_offset", C.c_uint64)]


class _CAlign(C.Structure):
    _fields_ = [("cigar", C.c_void_p), ("md", C.c_void_p), ("extended", C.c_void_p), ("position_offset", C.c_int32),
                ("qstart", C.c_int32), ("qend", C.c_int32), ("score", C.c_float), ("identity", C.c_float), ("nm", C.c_int32)]


ALIGN_REC = np.dtype([("position_offset", "<i4"), ("qstart", "<i4"), ("qend", "<i4"), ("nm", "<i4"), ("identity", "<f4"),
                      ("score", "<f4"), ("str_off", "<u4"), ("cigar_len", "<u2"), ("md_len", "<u2")])
PAIR = np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])
# ngm_b200_align == the reference's struct Align (IAlignment.h:14-28) on LP64
ALIGN_C = np.dtype([("cigar", "<u8"), ("md", "<u8"), ("extended", "<u8"), ("position_offset", "<i4"), ("qstart", "<i4"), ("qend", "<i4"),
                    ("score", "<f4"), ("identity", "<f4"), ("nm", "<i4")])
assert ALIGN_C.itemsize == C.sizeof(_CAlign)
PF_REVERSE, PF_DIR, PF_SKIP = 1, 2, 4


@dataclass
class Align:
    """The reference's ``struct Align`` (IAlignment.h:14-28)."""
    pBuffer1: bytes      # CIGAR
    pBuffer2: bytes      # MD
    PositionOffset: int
    QStart: int
    QEnd: int
    Score: float
    Identity: float
    NM: int


_lib = None


def load_library() -> C.CDLL:
    """Load the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NgmB200Error(f"{LIB_PATH} not built: run `python -m nextgenmap_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    lib.ngm_b200_last_error.restype = C.c_char_p
    lib.ngm_b200_create.restype = C.c_void_p
    lib.ngm_b200_create.argtypes = [C.POINTER(Params)]
    lib.ngm_b200_destroy.argtypes = [C.c_void_p]
    lib.ngm_b200_launch_count.restype = C.c_uint64
    lib.ngm_b200_launch_count.argtypes = [C.c_void_p]
    for name in ("ngm_b200_score_batch_size", "ngm_b200_align_batch_size"):
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.ngm_b200_batch_score.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_batch_align.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.ngm_b200_set_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.ngm_b200_score_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ngm_b200_align_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ngm_b200_dev_set_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.ngm_b200_dev_gather_winners.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_set_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.ngm_b200_dev_score_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_align_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_gather_winners_scored.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_align_pairs_scored.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.ngm_b200_dev_select_top1_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_cs_build_index.argtypes = [C.c_void_p, C.POINTER(CsParams), C.c_void_p, C.c_uint32]
    lib.ngm_b200_cs_load_index.argtypes = [C.c_void_p, C.POINTER(CsParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    lib.ngm_b200_cs_index_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
    lib.ngm_b200_cs_export_index.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_cs_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.c_void_p]
    lib.ngm_b200_cs_exact_reads.restype = C.c_uint64
    lib.ngm_b200_cs_exact_reads.argtypes = [C.c_void_p]
    lib.ngm_b200_dev_select_topn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
    lib.ngm_b200_map_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(MapResult)]
    lib.ngm_b200_format_sam.argtypes = [C.c_void_p, C.POINTER(SamOpts), C.POINTER(SamBatch), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.ngm_b200_pe_configure.argtypes = [C.c_void_p, C.POINTER(PeParams)]
    lib.ngm_b200_pe_insert_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ngm_b200_pe_deferred_fragments.argtypes = [C.c_void_p]
    lib.ngm_b200_pe_deferred_fragments.restype = C.c_int64
    lib.ngm_b200_dev_select_pairs.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p]
    lib.ngm_b200_cs_estimate_sensitivity.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.ngm_b200_cs_set_sensitivity.argtypes = [C.c_void_p, C.c_float]
    lib.ngm_b200_cs_configure_mutation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ngm_b200_dev_cs_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.c_void_p]
    lib.ngm_b200_read_ht_file.argtypes = [C.c_char_p, C.POINTER(_CHtFile)]
    lib.ngm_b200_free_ht_file.argtypes = [C.POINTER(_CHtFile)]
    lib.ngm_b200_write_ht_file.argtypes = [C.c_char_p, C.POINTER(_CHtFile)]
    lib.ngm_b200_dev_select_top1.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ngm_b200_run_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut)]
    lib.ngm_b200_dev_run_batch.argtypes = [C.c_void_p, C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_void_p]
    lib.ngm_b200_set_pipeline.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.ngm_b200_se_configure.argtypes = [C.c_void_p, C.c_int]
    lib.ngm_b200_se_configure_topn.argtypes = [C.c_void_p, C.c_int]
    lib.ngm_b200_host_alloc.restype = C.c_void_p
    lib.ngm_b200_host_alloc.argtypes = [C.c_size_t]
    lib.ngm_b200_host_free.argtypes = [C.c_void_p]
    lib.ngm_b200_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.ngm_b200_host_unregister.argtypes = [C.c_void_p]
    lib.ngm_b200_pack_reads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    _lib = lib
    return lib


def _row_pointers(a: np.ndarray) -> np.ndarray:
    """char** for a C-contiguous uint8 [n, w] array."""
    return a.ctypes.data + np.arange(a.shape[0], dtype=np.uint64) * np.uint64(a.strides[0])


class CudaSW:
    """B200 implementation of the reference's IAlignment plugin (SWOclCigar equivalent)."""

    def __init__(self, qry_max_len: int, corridor: int, match_bonus: float = 10, mismatch_penalty: float = 15,
                 gap_read_penalty: float = 20, gap_ref_penalty: float = 20, match_bonus_tt: float = 0,
                 match_bonus_tc: float = 0, bs_mapping: int = 0, slam_seq: int = 0, hard_clip: int = 0,
                 silent_clip: int = 0, device: int = 0, score_batch: int = 0, align_batch: int = 0, lane_mode: int = 0):
        self.lib = load_library()
        self.params = Params(qry_max_len, corridor, match_bonus, mismatch_penalty, gap_read_penalty, gap_ref_penalty,
                             match_bonus_tt, match_bonus_tc, bs_mapping, slam_seq, hard_clip, silent_clip, device,
                             score_batch, align_batch, lane_mode)
        self.qml, self.corridor = qry_max_len, corridor
        self.ctx = self.lib.ngm_b200_create(C.byref(self.params))
        if not self.ctx:
            raise NgmB200Error(self._err())

    def _err(self) -> str:
        return self.lib.ngm_b200_last_error().decode(errors="replace")

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise NgmB200Error(f"ngm_b200 error {rc}: {self._err()}")
        return rc

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.ngm_b200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- IAlignment ---------------------------------------------------------
    def GetScoreBatchSize(self) -> int:
        return self.lib.ngm_b200_score_batch_size(self.ctx)

    def GetAlignBatchSize(self) -> int:
        return self.lib.ngm_b200_align_batch_size(self.ctx)

    def launch_count(self) -> int:
        return int(self.lib.ngm_b200_launch_count(self.ctx))

    def _rows(self, refSeqList, qrySeqList):
        refs = np.ascontiguousarray(refSeqList, dtype=np.uint8)
        qrys = np.ascontiguousarray(qrySeqList, dtype=np.uint8)
        if refs.ndim != 2 or qrys.ndim != 2 or refs.shape[0] != qrys.shape[0]:
            raise ValueError("refSeqList / qrySeqList must be [n, width] byte arrays of equal n")
        if refs.shape[1] < self.qml + self.corridor or qrys.shape[1] < self.qml:
            raise ValueError("rows shorter than qry_max_len+corridor / qry_max_len")
        return refs, qrys

    def BatchScore(self, mode: int, refSeqList, qrySeqList, extData: Optional[Sequence[int]] = None) -> np.ndarray:
        """IAlignment::BatchScore -> float32 scores (SWOcl.cpp:33-162)."""
        refs, qrys = self._rows(refSeqList, qrySeqList)
        n = refs.shape[0]
        out = np.full(n, np.nan, dtype=np.float32)
        if n == 0:
            return out
        rp, qp = _row_pointers(refs), _row_pointers(qrys)
        d = None if extData is None else np.ascontiguousarray(extData, dtype=np.uint8)
        got = self._check(self.lib.ngm_b200_batch_score(self.ctx, mode, n, rp.ctypes.data, qp.ctypes.data, out.ctypes.data,
                                                        None if d is None else d.ctypes.data))
        assert got == n
        return out

    def BatchAlign(self, mode: int, refSeqList, qrySeqList, qalSeqList=None, extData: Optional[Sequence[int]] = None) -> List[Align]:
        """IAlignment::BatchAlign -> list of Align (SWOclCigar.cpp:104-370)."""
        refs, qrys = self._rows(refSeqList, qrySeqList)
        n = refs.shape[0]
        if n == 0:
            return []
        res, cig, md = self.batch_align_raw(mode, refs, qrys, extData)
        out = []
        for i in range(n):
            r = res[i]
            out.append(Align(cig[i].tobytes().split(b"\0")[0], md[i].tobytes().split(b"\0")[0], int(r["position_offset"]), int(r["qstart"]),
                             int(r["qend"]), float(r["score"]), float(r["identity"]), int(r["nm"])))
        return out

    def alloc_align_buffers(self, n: int):
        """What AlignmentBuffer allocates ONCE per thread (AlignmentBuffer.h:63-90, AlignmentBuffer.cpp:106-109): the struct Align array and the
        CIGAR / MD rows of 4*qry_max_len+ bytes its pBuffer1 / pBuffer2 point to.  -> (res, cig, md), touched, for ``batch_align_raw(buffers=)``."""
        stride = 4 * max(1, self.qml) + 2 * self.corridor + 64
        cig = np.zeros((n, stride), np.uint8)
        md = np.zeros((n, stride), np.uint8)
        res = np.zeros(n, dtype=ALIGN_C)
        res["cigar"] = _row_pointers(cig)
        res["md"] = _row_pointers(md)
        return res, cig, md

    def batch_align_raw(self, mode: int, refs: np.ndarray, qrys: np.ndarray, extData=None, buffers=None):
        """BatchAlign into numpy buffers: (struct Align array, CIGAR rows, MD rows); rows are 4*qry_max_len+ bytes
        pre-filled with "!!!" like AlignmentBuffer.cpp:108-109.  buffers: the result of ``alloc_align_buffers`` (reused by a caller that,
        like AlignmentBuffer, keeps its result buffers over the batches)."""
        n = refs.shape[0]
        res, cig, md = buffers if buffers is not None else self.alloc_align_buffers(n)
        res, cig, md = res[:n], cig[:n], md[:n]
        cig[:, :3] = 0x21
        md[:, :3] = 0x21
        cig[:, 3] = 0
        md[:, 3] = 0
        rp, qp = _row_pointers(refs), _row_pointers(qrys)
        d = None if extData is None else np.ascontiguousarray(extData, dtype=np.uint8)
        got = self._check(self.lib.ngm_b200_batch_align(self.ctx, mode, n, rp.ctypes.data, qp.ctypes.data, qp.ctypes.data,
                                                        res.ctypes.data, None if d is None else d.ctypes.data))
        assert got == n
        return res, cig, md

    # -- descriptor fast path -----------------------------------------------
    def set_reference(self, packed: np.ndarray, concat_len: int) -> None:
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self._check(self.lib.ngm_b200_set_reference(self.ctx, packed.ctypes.data, concat_len))

    def set_reads(self, reads: np.ndarray) -> None:
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        self._check(self.lib.ngm_b200_set_reads(self.ctx, reads.ctypes.data, reads.shape[0], reads.shape[1]))

    def score_pairs(self, mode: int, pairs: np.ndarray) -> np.ndarray:
        pairs = np.ascontiguousarray(pairs, dtype=PAIR)
        out = np.full(len(pairs), np.nan, dtype=np.float32)
        if len(pairs):
            self._check(self.lib.ngm_b200_score_pairs(self.ctx, mode, len(pairs), pairs.ctypes.data, out.ctypes.data))
        return out

    def align_pairs(self, mode: int, pairs: np.ndarray, str_capacity: Optional[int] = None):
        pairs = np.ascontiguousarray(pairs, dtype=PAIR)
        n = len(pairs)
        recs = np.zeros(n, dtype=ALIGN_REC)
        cap = str_capacity or max(4096, 64 * n)
        used = C.c_size_t(0)
        for _ in range(2):
            heap = np.zeros(cap, dtype=np.uint8)
            rc = self.lib.ngm_b200_align_pairs(self.ctx, mode, n, pairs.ctypes.data, recs.ctypes.data, heap.ctypes.data, cap, C.byref(used))
            if rc == -3 and used.value > cap:
                cap = used.value + 64
                continue
            self._check(rc)
            break
        return recs, heap[: used.value]

    # -- whole batches: ScoreBuffer::DoRun + AlignmentBuffer::DoRun in one pipelined call --------------
    def set_pipeline(self, lanes: int = 3, sub_batch_reads: int = 1 << 20) -> None:
        self._check(self.lib.ngm_b200_set_pipeline(self.ctx, lanes, sub_batch_reads))

    def se_configure(self, strata: int = 0, topn: int = 1) -> None:
        """`strata` and `topn` (NGM --strata / -n) of the single-end batch entry points; topn > 1: ScoreBuffer::topNSE."""
        self._check(self.lib.ngm_b200_se_configure(self.ctx, strata))
        self._check(self.lib.ngm_b200_se_configure_topn(self.ctx, topn))
        self.se_topn = topn

    def pack_reads(self, reads: np.ndarray, threads: int = 0):
        """ASCII rows -> (packed2 uint8 [n, row_bytes], read_len uint16 [n], exceptions READ_EXC [k]) on the host (ngm_b200_pack_reads)."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, stride = reads.shape
        row_bytes = 4 * ((stride + 15) // 16)
        packed = np.zeros((n, row_bytes), np.uint8)
        lens = np.zeros(n, np.uint16)
        cap = 1024
        need = C.c_size_t(0)
        for _ in range(2):
            exc = np.zeros(cap, dtype=READ_EXC)
            rc = self.lib.ngm_b200_pack_reads(reads.ctypes.data, n, stride, packed.ctypes.data, row_bytes, lens.ctypes.data, exc.ctypes.data, cap, C.byref(need), threads)
            if rc == -3 and need.value > cap:
                cap = need.value
                continue
            self._check(rc)
            break
        return packed, lens, exc[: need.value]

    def run_batch(self, mode: int, reads: np.ndarray, cand_begin: np.ndarray, pairs: np.ndarray, paired: bool = False, packed: bool = False,
                  desc_u64: bool = False, str_capacity: Optional[int] = None, want_scores: bool = True) -> dict:
        """ngm_b200_run_batch on host arrays.  reads: ASCII rows; packed=True sends them 2-bit packed (ngm_b200_pack_reads);
        pairs: PAIR records (read_index is ignored); desc_u64=True sends 64-bit descriptors.  After se_configure(topn > 1) a single-end
        batch returns recs [n, topn], sel [n, topn] and n_sel [n]."""
        topn = 1 if paired else getattr(self, "se_topn", 1)
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        cand_begin = np.ascontiguousarray(cand_begin, dtype=np.int32)
        pairs = np.ascontiguousarray(pairs, dtype=PAIR)
        n = reads.shape[0]
        npairs = int(cand_begin[-1] - cand_begin[0])
        keep = []
        bi = BatchIn()
        bi.n_reads, bi.mode, bi.paired = n, mode, 1 if paired else 0
        if packed:
            pk, lens, exc = self.pack_reads(reads)
            keep += [pk, lens, exc]
            bi.read_format, bi.reads, bi.read_stride = READS_PACKED2, pk.ctypes.data, pk.shape[1]
            bi.read_len, bi.exceptions, bi.n_exceptions = lens.ctypes.data, (exc.ctypes.data if len(exc) else None), len(exc)
        else:
            bi.read_format, bi.reads, bi.read_stride = READS_ASCII, reads.ctypes.data, reads.shape[1]
        if desc_u64:
            d = np.ascontiguousarray(make_desc_u64(pairs["window_start"], pairs["flags"]))
            keep.append(d)
            bi.desc_format, bi.desc = DESC_U64, d.ctypes.data
        else:
            bi.desc_format, bi.desc = DESC_PAIR16, pairs.ctypes.data
        bi.cand_begin = cand_begin.ctypes.data
        res = {"scores": np.full(max(npairs, 1), np.nan, np.float32), "best_pair": np.zeros(n, np.int32), "mapq": np.zeros(n, np.int32),
               "num_top": np.zeros(n, np.int32), "pair_fail": np.zeros(n, np.int32), "recs": np.zeros((n, topn) if topn > 1 else n, dtype=ALIGN_REC)}
        if topn > 1:
            res["sel"], res["n_sel"] = np.zeros((n, topn), np.int32), np.zeros(n, np.int32)
        cap = str_capacity if str_capacity is not None else max(4096, 96 * n * topn)
        for _ in range(2):
            heap = np.zeros(max(cap, 1), np.uint8)
            bo = BatchOut(res["scores"].ctypes.data if want_scores else None, res["best_pair"].ctypes.data, res["mapq"].ctypes.data, res["num_top"].ctypes.data,
                          res["pair_fail"].ctypes.data if paired else None, res["recs"].ctypes.data, heap.ctypes.data, cap, 0, None,
                          res["sel"].ctypes.data if topn > 1 else None, res["n_sel"].ctypes.data if topn > 1 else None)
            rc = self.lib.ngm_b200_run_batch(self.ctx, C.byref(bi), C.byref(bo))
            if rc == -3 and bo.str_used > cap:
                cap = int(bo.str_used)
                continue
            self._check(rc)
            break
        res["scores"] = res["scores"][:npairs]
        res["heap"] = heap
        res["str_used"] = int(bo.str_used)
        return res

    # -- candidate search (CS / CompactPrefixTable) ---------------
    @staticmethod
    def cs_params(kmer: int = 13, kmer_skip: int = 2, bin_size: int = 2, skip_rep: int = 1, sensitivity: float = 0.5, kmer_min: float = 0.0,
                  max_kfreq: int = 0, max_cmrs: int = 0) -> CsParams:
        return CsParams(kmer, kmer_skip, bin_size, skip_rep, sensitivity, kmer_min, max_kfreq, max_cmrs)

    def cs_build_index(self, contigs, params: Optional[CsParams] = None) -> dict:
        """contigs: [(start, length)] in concatenated coordinates (SeqStart, SeqLen).  Needs set_reference first."""
        params = params or self.cs_params()
        arr = (_CContigRec * len(contigs))()
        for i, (st, ln) in enumerate(contigs):
            arr[i].start, arr[i].length, arr[i].name_len = st, ln, 0
        self._check(self.lib.ngm_b200_cs_build_index(self.ctx, C.byref(params), arr, len(contigs)))
        return self.cs_index_info()

    def cs_load_index(self, tab: np.ndarray, weight: np.ndarray, table: np.ndarray, params: Optional[CsParams] = None) -> dict:
        params = params or self.cs_params()
        tab = np.ascontiguousarray(tab, dtype=np.uint32)
        weight = np.ascontiguousarray(weight, dtype=np.int8)
        table = np.ascontiguousarray(table, dtype=np.uint32)
        self._check(self.lib.ngm_b200_cs_load_index(self.ctx, C.byref(params), tab.ctypes.data, weight.ctypes.data, len(tab), table.ctypes.data, len(table)))
        return self.cs_index_info()

    def cs_index_info(self) -> dict:
        il, tl, mk = C.c_uint32(0), C.c_uint32(0), C.c_int32(0)
        self._check(self.lib.ngm_b200_cs_index_info(self.ctx, C.byref(il), C.byref(tl), C.byref(mk)))
        return {"index_len": il.value, "table_len": tl.value, "max_kfreq": mk.value}

    def cs_export_index(self):
        info = self.cs_index_info()
        tab = np.zeros(info["index_len"], np.uint32)
        weight = np.zeros(info["index_len"], np.int8)
        table = np.zeros(max(info["table_len"], 1), np.uint32)
        self._check(self.lib.ngm_b200_cs_export_index(self.ctx, tab.ctypes.data, weight.ctypes.data, table.ctypes.data))
        return tab, weight, table[: info["table_len"]]

    def cs_search(self, reads: np.ndarray, exact_only: bool = False, capacity: Optional[int] = None):
        """-> (cand_begin int32 [n+1], pairs PAIR [total], votes float32 [total], max_hit float32 [n])"""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n, stride = reads.shape
        begin = np.zeros(n + 1, np.int32)
        mh = np.zeros(n, np.float32)
        cap = capacity or max(1024, 8 * n)
        total = C.c_size_t(0)
        for _ in range(2):
            pairs = np.zeros(cap, dtype=PAIR)
            votes = np.zeros(cap, np.float32)
            rc = self.lib.ngm_b200_cs_search(self.ctx, reads.ctypes.data, n, stride, 1 if exact_only else 0, begin.ctypes.data, pairs.ctypes.data,
                                             votes.ctypes.data, cap, C.byref(total), mh.ctypes.data)
            if rc == -3 and total.value > cap:
                cap = total.value + 16
                continue
            self._check(rc)
            break
        return begin, pairs[: total.value], votes[: total.value], mh

    def cs_estimate_sensitivity(self, reads: np.ndarray, install: bool = True) -> float:
        """ReadProvider::init's estimate from the whole read set (every 1000th read, ReadProvider.cpp:240); 0.5 for < 1000 reads."""
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        n = reads.shape[0]
        sens = C.c_float(0.5)
        if n >= 1000:
            sample = np.ascontiguousarray(reads[999:min(n, 10_000_000 - 1):1000])
            self._check(self.lib.ngm_b200_cs_estimate_sensitivity(self.ctx, sample.ctypes.data, sample.shape[0], sample.shape[1], C.byref(sens)))
        if install:
            self._check(self.lib.ngm_b200_cs_set_sensitivity(self.ctx, sens))
        return float(sens.value)

    def cs_configure_mutation(self, bs_mapping: int = 0, slam_seq: int = 0, bs_cutoff: int = 6, paired: bool = False, read_kmer_skip: int = 2) -> None:
        """CS::RunBatch under --bs-mapping / --slam-seq 4|x (CS::PrefixMutateSearch, CS.cpp:53-112,341-380); both 0 = off."""
        self._check(self.lib.ngm_b200_cs_configure_mutation(self.ctx, bs_mapping, slam_seq, bs_cutoff, 1 if paired else 0, read_kmer_skip))

    def pe_configure(self, pair_score_cutoff: float = 0.9, min_insert_size: int = 0, max_insert_size: int = 1000, strata: int = 0,
                     fast_pairing: int = 0) -> None:
        """Paired-end selection parameters (Config.cpp:393-406); resets the running insert-size sums of ScoreBuffer (ScoreBuffer.h:90)."""
        p = PeParams(pair_score_cutoff, min_insert_size, max_insert_size, strata, fast_pairing)
        self._check(self.lib.ngm_b200_pe_configure(self.ctx, C.byref(p)))

    def pe_insert_stats(self):
        """-> (pairDistSum, pairDistCount) of ScoreBuffer (ScoreBuffer.cpp:420-422)."""
        s, n = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.ngm_b200_pe_insert_stats(self.ctx, C.byref(s), C.byref(n)))
        return int(s.value), int(n.value)

    def pe_deferred_fragments(self) -> int:
        return int(self.lib.ngm_b200_pe_deferred_fragments(self.ctx))

    def cs_exact_reads(self) -> int:
        return int(self.lib.ngm_b200_cs_exact_reads(self.ctx))

    def cs_exact_reasons(self) -> dict:
        names = ["hits", "wrap", "queue", "table", "multi", "zero_threshold", "accepted", "items", "order"]
        out = (C.c_uint32 * len(names))()
        self.lib.ngm_b200_cs_exact_reasons.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._check(self.lib.ngm_b200_cs_exact_reasons(self.ctx, out, len(names)))
        return {k: int(v) for k, v in zip(names, out)}

    @staticmethod
    def strings_of(recs: np.ndarray, heap: np.ndarray, i: int):
        r = recs[i]
        o = int(r["str_off"])
        raw = heap.tobytes()
        return raw[o: o + int(r["cigar_len"])], raw[o + int(r["cigar_len"]): o + int(r["cigar_len"]) + int(r["md_len"])]


# -- NGM's encoded-reference cache file --------------------------------------------------
class _CContig(C.Structure):
    _fields_ = [("start", C.c_uint64), ("length", C.c_uint32), ("name_len", C.c_uint32), ("name", C.c_char * 100)]


class _CEncRef(C.Structure):
    _fields_ = [("concat_len", C.c_uint64), ("packed_bytes", C.c_uint64), ("n_contigs", C.c_uint32), ("packed", C.POINTER(C.c_uint8)),
                ("contigs", C.POINTER(_CContig))]


class EncodedReference:
    """`<ref>-enc.2.ngm` (SequenceProvider.cpp:189-208) + convert() (SequenceProvider.cpp:111-141).  Host only."""

    def __init__(self, path: str):
        self.lib = load_library()
        self.lib.ngm_b200_read_enc_ref.argtypes = [C.c_char_p, C.POINTER(_CEncRef)]
        self.lib.ngm_b200_free_enc_ref.argtypes = [C.POINTER(_CEncRef)]
        self.lib.ngm_b200_convert.argtypes = [C.POINTER(_CEncRef), C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        self.c = _CEncRef()
        rc = self.lib.ngm_b200_read_enc_ref(str(path).encode(), C.byref(self.c))
        if rc != 0:
            raise NgmB200Error(f"cannot read encoded reference {path} (rc {rc})")
        self.concat_len = int(self.c.concat_len)
        self.packed = np.ctypeslib.as_array(self.c.packed, shape=(int(self.c.packed_bytes),)).copy()
        self.contigs = [(self.c.contigs[i].name[: self.c.contigs[i].name_len].decode(), int(self.c.contigs[i].start), int(self.c.contigs[i].length))
                        for i in range(int(self.c.n_contigs))]

    def convert(self, concat_pos: int):
        contig, pos = C.c_uint32(0), C.c_uint64(0)
        ok = self.lib.ngm_b200_convert(C.byref(self.c), concat_pos, C.byref(contig), C.byref(pos))
        return (int(contig.value), int(pos.value)) if ok else None

    def write(self, path: str) -> None:
        """writeEncRefToFile (SequenceProvider.cpp:189-208)."""
        self.lib.ngm_b200_write_enc_ref.argtypes = [C.c_char_p, C.POINTER(_CEncRef)]
        if self.lib.ngm_b200_write_enc_ref(str(path).encode(), C.byref(self.c)) != 0:
            raise NgmB200Error(f"cannot write {path}")

    def close(self):
        if self.c.packed:
            self.lib.ngm_b200_free_enc_ref(C.byref(self.c))


class PrefixTableFile:
    """`<ref>-ht-<k>-<skip>.3.ngm` (CompactPrefixTable::saveToFile / readFromFile, PrefixTable.cpp:819-921).  Host only."""

    def __init__(self, path: str):
        self.lib = load_library()
        c = _CHtFile()
        rc = self.lib.ngm_b200_read_ht_file(str(path).encode(), C.byref(c))
        if rc != 0:
            raise NgmB200Error(f"cannot read prefix table file {path} (rc {rc})")
        self.kmer, self.kmer_skip, self.index_len, self.table_len = int(c.kmer), int(c.kmer_skip), int(c.index_len), int(c.table_len)
        self.unit_offset = int(c.unit_offset)
        self.tab = np.ctypeslib.as_array(c.tab, shape=(self.index_len,)).copy()
        self.weight = np.ctypeslib.as_array(c.weight, shape=(self.index_len,)).copy()
        self.table = np.ctypeslib.as_array(c.table, shape=(max(self.table_len, 1),)).copy()[: self.table_len]
        self.lib.ngm_b200_free_ht_file(C.byref(c))

    @staticmethod
    def write(path: str, kmer: int, kmer_skip: int, tab: np.ndarray, weight: np.ndarray, table: np.ndarray) -> None:
        lib = load_library()
        tab = np.ascontiguousarray(tab, dtype=np.uint32)
        weight = np.ascontiguousarray(weight, dtype=np.int8)
        table = np.ascontiguousarray(table, dtype=np.uint32)
        c = _CHtFile(kmer, kmer_skip, len(tab), len(table), tab.ctypes.data_as(C.POINTER(C.c_uint32)), weight.ctypes.data_as(C.POINTER(C.c_int8)),
                     table.ctypes.data_as(C.POINTER(C.c_uint32)), 0)
        if lib.ngm_b200_write_ht_file(str(path).encode(), C.byref(c)) != 0:
            raise NgmB200Error(f"cannot write {path}")
