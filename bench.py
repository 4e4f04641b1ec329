#!/usr/bin/env python
"""bench.py -- reads/s of NGM's seed-and-extend hot path (BatchScore -> top-1 -> BatchAlign).

Workload (BASELINE.json configs[1]): 10 M x 150 bp single-end synthetic reads against a 3 Gbp
synthetic reference (24 x 125 Mbp contigs), ~1.5 candidate windows per read, local alignment,
default scoring (10/-15/-20/-20), qry_max_len 152, corridor 27.  One "step" = one pass of the hot
path over the whole read set of this rank:

    pack reads (+ reverse complements) -> score every (read, window) pair -> top-1 selection + MAPQ
    -> banded alignment with backtrace of the winner -> CIGAR / MD / NM / identity records

* ``value``  : device-resident throughput -- reads, descriptors and the packed reference are in
               HBM when the timed region starts; results stay in HBM.
* ``e2e``    : the same pipeline through the C ABI with HOST (pinned) inputs and outputs: every
               step copies reads + descriptors host->device and records + strings device->host,
               sub-batches ping-pong over two streams.
* ``--impl reference`` : the reference's own CPU implementation of the path (oracle/_ref: its
               unmodified OpenCL kernels on the vendored AMD CPU runtime, one instance per host
               core like ``ngm -t``), on a bounded sample of the same workload.

Multi-GPU (torchrun): reads shard across ranks (weak scaling: every rank owns a full-size read
set), each rank holds the whole reference; NCCL only broadcasts the packed reference at start-up
and reduces the mapping counters at the end.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

READ_LEN = 150
MODE_LOCAL = 0
ALG_BYTES_SCORE = 180          # ceil(L/2) + ceil((L+corridor)/2) + 12 + 4 at 150 bp
CELLS_PER_PAIR = 150 * 27      # L x corridor


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("NGM_BENCH_READS", 10_000_000)))
    ap.add_argument("--contigs", type=int, default=int(os.environ.get("NGM_BENCH_CONTIGS", 24)))
    ap.add_argument("--contig-len", type=int, default=int(os.environ.get("NGM_BENCH_CONTIG_LEN", 125_000_000)))
    ap.add_argument("--sub-batch", type=int, default=524_288, help="reads per sub-batch inside ngm_b200_run_batch / ngm_b200_map_batch")
    ap.add_argument("--lanes", type=int, default=4, help="lanes (stream + staging) the sub-batches rotate over inside the library")
    ap.add_argument("--read-len", type=int, default=READ_LEN, help="read length (BASELINE configs[4] sweep: 75/100/150/250/400); default 150")
    ap.add_argument("--corridor", type=int, default=0, help="override the band width (NGM -C n => 2n); 0 = int(5 + 0.15 L)")
    ap.add_argument("--sub-rate", type=float, default=0.01)
    ap.add_argument("--indel-rate", type=float, default=0.0005)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cs", action="store_true", help="skip the candidate-search section (index build + k-mer vote on device)")
    ap.add_argument("--no-pe", action="store_true", help="skip the paired-end section (needs the candidate-search section)")
    ap.add_argument("--sensitivity", type=float, default=0.5, help="CS sensitivity (NGM -s; its own default when not estimated)")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[i]: 1 = 10 M x 150 bp SE (headline, default); 2 = paired-end 2 x 150 bp (10 M reads = 5 M fragments per GPU); "
                         "3 = 5 M x 250 bp, 12 %% substitutions + 1.5 %% ins + 1.5 %% del, -C 40 (corridor 80); 4 = read-length sweep point, use with --read-len")
    ap.add_argument("--no-ngm", action="store_true", help="skip the whole-program run of the unmodified ngm on the host cores (candidate_search.cpu_baseline_ngm)")
    ap.add_argument("--ngm-sample", type=int, default=1_000_000, help="reads of the workload the unmodified ngm maps (SURVEY 8d)")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: do not bind the rank to the CPUs next to its GPU before the pinned staging is allocated")
    args = ap.parse_args()
    if args.config == 3:
        args.read_len, args.corridor, args.sub_rate, args.indel_rate = 250, 80, 0.12, 0.03
        if "NGM_BENCH_READS" not in os.environ and "--reads" not in sys.argv:
            args.reads = 5_000_000
    return args


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # started a little before the timed region (nvidia-smi needs ~0.2 s to produce its first line); only
    # samples whose arrival time falls inside [mark_begin, mark_end] are used

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if self.t_begin is None or (self.t_begin <= t <= (self.t_end or t) + 0.15)]
        if not inside:
            inside = [ln for (_, ln) in self.lines[-3:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU baseline sample (shared by cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------
def cpu_sample_windows(packed_host: np.ndarray, concat_len: int, reads: np.ndarray, pairs: np.ndarray, cand_begin: np.ndarray,
                       n_sample: int, qml: int, corridor: int):
    """Decode the sample's windows the way ScoreBuffer does (ScoreBuffer.cpp:113-118) and order the pairs
    so that each read's first candidate (the true locus) comes first: those are the ones aligned."""
    from oracle import port
    buf_len = ((qml + corridor) | 1) + 1
    comp = np.arange(256, dtype=np.uint8)
    for a, b in zip(b"ACGT", b"TGCA"):
        comp[a] = b
    first, rest = [], []
    for r in range(n_sample):
        for k, j in enumerate(range(cand_begin[r], cand_begin[r + 1])):
            (first if k == 0 else rest).append(j)
    order = first + rest
    refs = np.zeros((len(order), buf_len), np.uint8)
    qrys = np.zeros((len(order), qml), np.uint8)
    for i, j in enumerate(order):
        start = int(pairs["window_start"][j])
        w = port.decode_window(packed_host, concat_len, start, buf_len)
        refs[i] = np.frombuffer(w, np.uint8) if w is not None else ord("N")
        row = reads[int(pairs["read_index"][j])]
        if int(pairs["flags"][j]) & 1:
            L = int(np.count_nonzero(row))
            rc = np.zeros_like(row)
            rc[:L] = comp[row[:L][::-1]]
            row = rc
        qrys[i] = row
    return refs, qrys, len(first)


def run_cpu_reference(refs, qrys, n_align, n_reads, qml, corridor, threads, warm, passes):
    """reads/s of the reference's own CPU path on the sample; falls back to the C port if oracle/_ref cannot run."""
    from oracle import ref_driver as rd
    if rd.available():
        try:
            r = rd.bench(refs, qrys, qml, corridor, MODE_LOCAL, n_align, threads, warm, passes)
            return {"value": n_reads * r["passes"] / r["wall_seconds"], "kind": "reference", "cores": threads,
                    "score_pairs_per_s": r["scored"] / max(r["score_seconds"], 1e-9), "align_pairs_per_s": r["aligned"] / max(r["align_seconds"], 1e-9),
                    "wall_seconds": r["wall_seconds"]}
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"[bench] oracle/_ref harness failed ({e}); timing the C port instead\n")
    from oracle import port
    t0 = time.perf_counter()
    for _ in range(passes):
        port.batch_score(refs, qrys, qml, corridor, MODE_LOCAL)
        port.batch_align(refs[:n_align], qrys[:n_align], qml, corridor, MODE_LOCAL)
    dt = time.perf_counter() - t0
    return {"value": n_reads * passes / dt, "kind": "port", "cores": 1, "wall_seconds": dt}


def bind_to_gpu_numa_node(torch, local_rank: int):
    """Pinned staging is first-touch: allocate it from the CPUs next to the rank's GPU (N > 1).  -> dict describing what was done."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        node = int((base / "numa_node").read_text().strip())
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if node < 0 or not use or use == allowed:
            return {"pci": bdf, "numa_node": node, "bound": False}
        os.sched_setaffinity(0, use)
        return {"pci": bdf, "numa_node": node, "bound": True, "cpus": len(use)}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "error": str(e)}


def run_ngm_whole(sw, ref, batch, n_sample, L, qml, contigs, sensitivity, threads, lib, tmp_root=None):
    """SURVEY 8d "CPU timing beside it": the UNMODIFIED NextGenMap (oracle/_ref/ngm/ngm_ref: its own CS, OpenCL-on-CPU kernels, selection,
    SAM writer) with `-t <threads>` on a sample of the same reads against the same 3 Gbp reference, with a warm index: the encoded
    reference and the prefix table are written in NGM's own cache-file formats (the table is the one built on the device, byte-identical
    to NGM's) so that NGM loads instead of rebuilding them.  -> reads/s from the wall clock of the mapping phase (a run with 1000 reads is
    subtracted as start-up: index load, OpenCL initialisation)."""
    import ctypes as C
    import re
    import shutil
    import tempfile
    from oracle import ngm_e2e as e2e
    from nextgenmap_b200.host.cuda_sw import PrefixTableFile, _CContig, _CEncRef
    if not e2e.available("ref"):
        return {"unavailable": "oracle/_ref/ngm/ngm_ref not built"}
    base = tmp_root or ("/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 12 << 30 else None)
    d = Path(tempfile.mkdtemp(prefix="ngm_whole_", dir=base))
    try:
        t0 = time.perf_counter()
        (d / "ref.fa").write_bytes(b">stub (NGM loads ref.fa-enc.2.ngm)\nACGT\n")
        packed = ref.packed.cpu().numpy()
        ctg = (_CContig * contigs)()
        for i, s0 in enumerate(ref.contig_start):
            nm = b"chr%d" % (i + 1)
            ctg[i].start, ctg[i].length, ctg[i].name_len, ctg[i].name = int(s0), int(ref.contig_len), len(nm), nm
        enc = _CEncRef()
        enc.concat_len, enc.packed_bytes, enc.n_contigs = ref.concat_len - 1, packed.size, contigs      # GetConcatRefLen() = binRefIndex - 1
        enc.packed = packed.ctypes.data_as(C.POINTER(C.c_uint8))
        enc.contigs = C.cast(ctg, type(enc.contigs))
        lib.ngm_b200_write_enc_ref.argtypes = [C.c_char_p, C.POINTER(_CEncRef)]
        if lib.ngm_b200_write_enc_ref(str(d / "ref.fa-enc.2.ngm").encode(), C.byref(enc)) != 0:
            return {"error": "cannot write the encoded reference"}
        tab, weight, table = sw.cs_export_index()
        PrefixTableFile.write(str(d / "ref.fa-ht-13-2.3.ngm"), 13, 2, tab, weight, table)
        del tab, weight, table, packed
        # FASTQ: fixed-width records, vectorised
        rd = batch.reads[:n_sample, :L].cpu().numpy()
        rec = np.empty((n_sample, 10 + L + 3 + L + 1), np.uint8)
        names = np.char.zfill(np.arange(n_sample).astype("S8"), 8)
        rec[:, 0] = ord("@")
        rec[:, 1:9] = np.frombuffer(names.tobytes(), np.uint8).reshape(n_sample, 8)
        rec[:, 9] = ord("\n")
        rec[:, 10:10 + L] = rd
        rec[:, 10 + L:13 + L] = np.frombuffer(b"\n+\n", np.uint8)
        rec[:, 13 + L:13 + 2 * L] = ord("I")
        rec[:, 13 + 2 * L] = ord("\n")
        (d / "reads.fq").write_bytes(rec.tobytes())
        (d / "warm.fq").write_bytes(rec[:1000].tobytes())
        prep_s = time.perf_counter() - t0
        env = dict(os.environ)
        ocl = e2e.HERE / "_ref" / "ocl"
        env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
        env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")

        def run(fq, out):
            t = time.perf_counter()
            p = subprocess.run([str(e2e.binary("ref")), "-r", str(d / "ref.fa"), "-q", str(d / fq), "-o", str(d / out), "-t", str(threads), "-s", str(sensitivity),
                                "--no-progress"], env=env, capture_output=True, text=True, cwd=d, timeout=900)
            wall = time.perf_counter() - t
            log = p.stdout + p.stderr
            m = re.search(r"Done \((\d+) reads mapped.*elapsed: ([0-9.]+)s", log)
            if p.returncode != 0 or m is None:
                raise RuntimeError("ngm failed: " + log[-600:])
            return wall, int(m.group(1)), float(m.group(2)), ("Reading encoded reference" in log), ("Reading RefTable" in log or "eading" in log)
        w0, _, e0, _, _ = run("warm.fq", "warm.sam")            # start-up: index load, OpenCL initialisation, 1000 reads
        w1, mapped, e1, enc_loaded, _ = run("reads.fq", "out.sam")
        map_s = max(w1 - w0, 1e-6)
        return {"value": (n_sample - 1000) / map_s, "unit": "reads/s", "cores": threads, "kind": "reference (unmodified ngm, whole program)",
                "sample": f"{n_sample} reads of the same workload, `ngm -t {threads} -s {sensitivity}`, warm index (NGM's own cache files: encoded reference + prefix table)",
                "wall_seconds": w1, "startup_wall_seconds": w0, "ngm_elapsed_seconds": e1, "ngm_startup_elapsed_seconds": e0, "mapped": mapped,
                "whole_run_reads_per_s_including_startup": n_sample / w1, "encoded_reference_loaded_from_cache": enc_loaded, "prepare_files_seconds": prep_s}
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---------------------------------------------------------------------------
_JSON_FD = None


def emit_json(line: dict) -> None:
    """The ONE line of stdout.  Everything else this process (or a library in it: NCCL's version banner) prints goes to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)            # keep the real stdout for the JSON line ...
    os.dup2(2, 1)                   # ... and send every other write to fd 1 (C libraries included) to stderr
    import torch
    import torch.distributed as dist

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # NCCL prints its version banner on stdout otherwise; stdout carries ONE JSON line
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    distributed = world > 1

    if args.impl == "reference" and rank != 0:
        return 0                                    # rank 0 alone runs the CPU arm

    from nextgenmap_b200 import workload
    L = args.read_len
    qml, corridor = workload.shapes_for(L)
    if args.corridor:
        corridor = args.corridor
    global ALG_BYTES_SCORE, CELLS_PER_PAIR
    ALG_BYTES_SCORE = (L + 1) // 2 + (L + corridor + 1) // 2 + 12 + 4
    CELLS_PER_PAIR = L * corridor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the workload is generated on the GPU; the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed and args.impl != "reference":
        dist.init_process_group("nccl", device_id=dev)
    numa = bind_to_gpu_numa_node(torch, local_rank) if (distributed and not args.no_numa and args.impl != "reference") else {"bound": False}

    # ---- workload ---------------------------------------------------------
    t_setup = time.perf_counter()
    from nextgenmap_b200 import sharding
    if distributed and args.impl != "reference":
        # rank 0 builds the reference, NCCL broadcasts the packed bytes over NVLink
        if rank == 0:
            ref = workload.make_reference(dev, args.contigs, args.contig_len, seed=20261017)
            packed, concat_len = sharding.broadcast_reference(ref.packed, ref.concat_len, src=0)
        else:
            packed, concat_len = sharding.broadcast_reference(torch.empty(0, dtype=torch.uint8, device=dev), 0, src=0)
            starts = workload.SPACER + np.arange(args.contigs, dtype=np.int64) * (args.contig_len + workload.SPACER)
            ref = workload.Reference(packed, concat_len, starts, args.contig_len)
    else:
        ref = workload.make_reference(dev, args.contigs, args.contig_len, seed=20261017)
    n_reads = args.reads
    if args.impl == "reference":
        n_reads = min(n_reads, 200_000)             # only a sample is needed on the CPU arm
    batch = workload.make_reads(ref, n_reads, L, qml, corridor, seed=20261018 + 1 + rank, sub_rate=args.sub_rate, indel_rate=args.indel_rate)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    cfg = {"workload": f"{args.reads} x {L} bp SE synthetic reads vs {args.contigs} x {args.contig_len} bp synthetic reference "
                       f"({ref.concat_len / 1e9:.2f} Gbp concatenated), ~1.5 candidates/read, local mode, qry_max_len {qml}, corridor {corridor}",
           "reads_per_step_per_gpu": args.reads, "pairs_per_step_per_gpu": batch.n_pairs if args.impl != "reference" else None,
           "step": "set_reads(pack+revcomp) -> score pairs -> top1+MAPQ -> gather winners -> align+backtrace+CIGAR/MD",
           "cache": "inputs larger than L2 (1.5 GB reads + 1.5 GB packed reference, random window gathers); no flush needed",
           "parallelism": f"read-sharded x{world}, reference replicated"}

    host_threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    # ---- reference arm ----------------------------------------------------
    if args.impl == "reference":
        n_sample = args.cpu_sample or min(n_reads, max(20_000, 3000 * host_threads))
        packed_host = ref.packed.cpu().numpy()
        reads_h = batch.reads[:n_sample].cpu().numpy()
        cb = batch.cand_begin[: n_sample + 1].cpu().numpy()
        pairs_h = batch.pairs[: int(cb[-1])].cpu().numpy().view(np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])).reshape(-1)
        refs, qrys, n_align = cpu_sample_windows(packed_host, ref.concat_len, reads_h, pairs_h, cb, n_sample, qml, corridor)
        res = run_cpu_reference(refs, qrys, n_align, n_sample, qml, corridor, host_threads, args.warmup, args.steps)
        sample = f"{n_sample} reads / {len(refs)} pairs of the workload per step, {res['cores']} host threads (one backend instance each, like ngm -t)"
        line = {"impl": "reference", "metric": "reads/sec aligned (150bp SE vs 3Gbp ref)", "value": res["value"], "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["wall_seconds"] / max(args.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (integer-valued)", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": res["kind"], "sample": sample},
                "e2e": {"value": res["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit_json(line)
        return 0

    # ---- B200 arm ---------------------------------------------------------
    import ctypes as C
    from nextgenmap_b200.host import CudaSW
    from nextgenmap_b200.host.cuda_sw import ALIGN_REC, BatchIn, BatchOut, READS_PACKED2, DESC_U64
    paired = args.config == 2
    if paired:                                       # BASELINE configs[2]: rows 2f / 2f + 1 are mates, insert ~ N(400, 40), FR
        batch = workload.make_reads(ref, n_reads, L, qml, corridor, seed=20261018 + 2 + 16 * rank, sub_rate=args.sub_rate, indel_rate=args.indel_rate, paired=True)
        cfg["workload"] = cfg["workload"].replace(" SE ", " paired-end (2 x, mates in adjacent rows) ")
    sw = CudaSW(qml, corridor, device=local_rank)
    lib, ctx = sw.lib, sw.ctx
    st = torch.cuda.current_stream().cuda_stream

    def check(rc):
        if rc < 0:
            raise RuntimeError(lib.ngm_b200_last_error().decode())
        return rc

    check(lib.ngm_b200_dev_set_reference(ctx, ref.packed.data_ptr(), ref.concat_len, st))
    if paired:
        sw.pe_configure()
    check(lib.ngm_b200_set_pipeline(ctx, max(1, args.lanes), min(args.sub_batch, max(2, batch.n_reads)) & ~1))
    n, npairs = batch.n_reads, batch.n_pairs
    # The caller's staging format (north_star: packed sequences in pinned staging buffers): reads 2 bit / base + lengths, one 64-bit
    # descriptor per (read, candidate), candidate offsets per read.  Packed on the host by the library's own ngm_b200_pack_reads.
    reads_np = batch.reads.cpu().numpy()
    sw.pack_reads(reads_np[: 1 << 16])
    t0 = time.perf_counter()
    pk_np, len_np, exc_np = sw.pack_reads(reads_np)            # (the call plus the wrapper's allocation of its three result arrays)
    pack_host_s = time.perf_counter() - t0
    del reads_np
    h_packed = torch.from_numpy(pk_np).pin_memory()
    h_len = torch.from_numpy(len_np.view(np.int16)).pin_memory()
    del pk_np, len_np
    ws = batch.pairs[:, 0:8].contiguous().view(torch.int64).reshape(-1)
    fl = batch.pairs[:, 12:16].contiguous().view(torch.int32).reshape(-1).long()
    lim = (1 << 56) - 1
    d_desc = (torch.where((ws < 0) | (ws > lim), torch.full_like(ws, lim), ws) | (fl << 56)).contiguous()      # NGM_B200_DESC()
    del ws, fl
    h_desc = d_desc.cpu().pin_memory()
    h_cb = batch.cand_begin.cpu().pin_memory()
    d_packed, d_len = h_packed.to(dev), h_len.to(dev)
    row_bytes = h_packed.shape[1]
    STR_PER = 48 + int(L * args.sub_rate * 5) + (64 if args.indel_rate > 0.001 else 0)      # string-heap bytes per read
    str_cap = STR_PER * n
    d_scores = torch.empty(max(npairs, 1), dtype=torch.float32, device=dev)
    d_best = torch.empty(n, dtype=torch.int32, device=dev)
    d_mapq = torch.empty(n, dtype=torch.int32, device=dev)
    d_ntop = torch.empty(n, dtype=torch.int32, device=dev)
    d_pfail = torch.empty(n, dtype=torch.int32, device=dev)
    d_recs = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    d_strings = torch.empty(str_cap, dtype=torch.uint8, device=dev)
    d_cursor = torch.zeros(1, dtype=torch.int32, device=dev)
    exc_ptr = exc_np.ctypes.data if len(exc_np) else None
    d_exc = torch.from_numpy(exc_np.view(np.uint8).reshape(-1, 8).copy()).to(dev) if len(exc_np) else None

    def batch_in(reads_ptr, len_ptr, exc_p, cb_ptr, desc_ptr):
        bi = BatchIn()
        bi.n_reads, bi.mode, bi.paired, bi.read_format = n, MODE_LOCAL, 1 if paired else 0, READS_PACKED2
        bi.reads, bi.read_stride, bi.desc_format, bi.read_len = reads_ptr, row_bytes, DESC_U64, len_ptr
        bi.exceptions, bi.n_exceptions, bi.n_desc, bi.cand_begin, bi.desc = exc_p, len(exc_np), npairs, cb_ptr, desc_ptr
        return bi

    dev_in = batch_in(d_packed.data_ptr(), d_len.data_ptr(), d_exc.data_ptr() if d_exc is not None else None, batch.cand_begin.data_ptr(), d_desc.data_ptr())
    dev_out = BatchOut(d_scores.data_ptr(), d_best.data_ptr(), d_mapq.data_ptr(), d_ntop.data_ptr(), d_pfail.data_ptr() if paired else None, d_recs.data_ptr(),
                       d_strings.data_ptr(), str_cap, 0, d_cursor.data_ptr())
    cfg["step"] = ("ngm_b200_dev_run_batch: expand 2-bit reads (+ reverse complements) -> resolve descriptors -> "
                   + ("BatchScore of every candidate -> top1PE (select_pairs) -> forward pass + backtrace + CIGAR/MD of every winner" if paired else
                      "forward pass with pointers over every candidate of reads with <= 4 candidates (BatchScore first for the others) -> top1 + MAPQ from the "
                      "forward maxima -> backtrace + CIGAR/MD of the winner"))

    def resident_step():
        check(lib.ngm_b200_dev_run_batch(ctx, C.byref(dev_in), C.byref(dev_out), st))

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        if paired:
            sw.pe_configure()
        resident_step()
    barrier()
    if int(d_cursor.item()) > str_cap:
        raise RuntimeError("string heap too small")

    launches0 = sw.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.mark_begin()
    e0.record()
    for _ in range(args.steps):
        resident_step()
    e1.record()
    barrier()
    clocks.mark_end()
    ms = e0.elapsed_time(e1)
    launches = sw.launch_count() - launches0
    clk = clocks.stop()
    if paired:                                       # leave the first-batch result in the buffers (the parity sample below starts from fresh sums)
        sw.pe_configure()
        resident_step()
        torch.cuda.synchronize()
    recs_all = d_recs.cpu().numpy().view(ALIGN_REC).reshape(-1)
    best_all = d_best.cpu().numpy()
    mapq_all = d_mapq.cpu().numpy()
    mapped = int(np.count_nonzero(recs_all["score"] >= 0))
    used_strings = int(d_cursor.item())
    heap_res = d_strings[:used_strings].cpu().numpy()

    # per-kernel timing on the launching stream (roofline): the classic device entry points on the same data
    def time_call(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    d_wpairs = torch.empty((n, 16), dtype=torch.uint8, device=dev)
    d_wscores = torch.empty(n, dtype=torch.float32, device=dev)
    d_scores_all = torch.empty(max(npairs, 1), dtype=torch.float32, device=dev)
    d_best1 = torch.empty(n, dtype=torch.int32, device=dev)
    d_mapq1 = torch.empty(n, dtype=torch.int32, device=dev)
    lib.ngm_b200_dev_set_reads_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    ms_expand = time_call(lambda: check(lib.ngm_b200_dev_set_reads_packed(ctx, d_packed.data_ptr(), n, row_bytes, d_len.data_ptr(),
                                                                       d_exc.data_ptr() if d_exc is not None else None, len(exc_np), st)))
    ms_pack = time_call(lambda: check(lib.ngm_b200_dev_set_reads(ctx, batch.reads.data_ptr(), n, qml, st)))
    ms_score = time_call(lambda: check(lib.ngm_b200_dev_score_pairs(ctx, MODE_LOCAL, npairs, batch.pairs.data_ptr(), d_scores_all.data_ptr(), st)))
    check(lib.ngm_b200_dev_select_top1(ctx, n, batch.cand_begin.data_ptr(), d_scores_all.data_ptr(), d_best1.data_ptr(), d_mapq1.data_ptr(), st))
    check(lib.ngm_b200_dev_gather_winners_scored(ctx, n, batch.pairs.data_ptr(), d_scores_all.data_ptr(), d_best1.data_ptr(), d_wpairs.data_ptr(), d_wscores.data_ptr(), st))
    d_recs1 = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    d_cursor1 = torch.zeros(1, dtype=torch.int32, device=dev)

    def align_only():
        d_cursor1.zero_()
        check(lib.ngm_b200_dev_align_pairs_scored(ctx, MODE_LOCAL, n, d_wpairs.data_ptr(), d_wscores.data_ptr(), d_recs1.data_ptr(), d_strings.data_ptr(),
                                                  str_cap, d_cursor1.data_ptr(), st))

    def align_unscored():
        d_cursor1.zero_()
        check(lib.ngm_b200_dev_align_pairs(ctx, MODE_LOCAL, n, d_wpairs.data_ptr(), d_recs1.data_ptr(), d_strings.data_ptr(), str_cap, d_cursor1.data_ptr(), st))

    ms_align_unscored = time_call(align_unscored, reps=2)
    ms_align = time_call(align_only)
    # forward / backtrace split of one align pass (events recorded inside the library between the two kernels of every launch set)
    lib.ngm_b200_profile.argtypes = [C.c_void_p, C.c_int]
    lib.ngm_b200_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.ngm_b200_alu_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    check(lib.ngm_b200_profile(ctx, 1))
    align_only()
    f_ms, b_ms = C.c_float(0), C.c_float(0)
    launch_sets = check(lib.ngm_b200_profile_read(ctx, C.byref(f_ms), C.byref(b_ms)))
    check(lib.ngm_b200_profile(ctx, 0))
    ms_fwd, ms_bt = float(f_ms.value), float(b_ms.value)
    # the same split inside one resident step (single-end: the forward pass runs over EVERY candidate of reads with <= 4 candidates, the
    # winner is picked from the forward maxima, only the winner is traced back)
    step_fwd_ms = step_bt_ms = None
    step_fwd_pairs = 0
    if not paired:
        check(lib.ngm_b200_profile(ctx, 1))
        resident_step()
        chunks = check(lib.ngm_b200_profile_read(ctx, C.byref(f_ms), C.byref(b_ms)))
        check(lib.ngm_b200_profile(ctx, 0))
        if chunks > 0:
            step_fwd_ms, step_bt_ms = float(f_ms.value), float(b_ms.value)
            cnt = batch.cand_begin[1:] - batch.cand_begin[:-1]
            # narrow local bands: forward pass over every candidate of reads with <= 4 candidates; wide bands (capacity > 48) score every
            # candidate first and run the forward pass (maximum known) on the winners only
            step_fwd_pairs = int(cnt[cnt <= 4].sum().item()) + int((cnt > 4).sum().item()) if corridor <= 48 else n
    pa, pi_, pm = C.c_double(0), C.c_double(0), C.c_double(0)
    check(lib.ngm_b200_alu_peak(ctx, C.byref(pa), C.byref(pi_), C.byref(pm)))
    alu_rate, imad_rate, mixed_rate = float(pa.value), float(pi_.value), float(pm.value)
    # the separate-call pipeline must agree with the fused batch (single-end: same winners, same records)
    classic_equal = None
    if not paired:
        r1 = d_recs1.cpu().numpy().view(ALIGN_REC).reshape(-1)
        classic_equal = bool(np.array_equal(d_best1.cpu().numpy(), best_all) and np.array_equal(d_mapq1.cpu().numpy(), mapq_all)
                             and all(np.array_equal(r1[f], recs_all[f]) for f in ["position_offset", "qstart", "qend", "nm", "identity", "score", "cigar_len", "md_len"])
                             and np.array_equal(d_scores_all.cpu().numpy(), d_scores.cpu().numpy()))
    del d_recs1

    # ---- end to end through the C ABI with host buffers: ONE ngm_b200_run_batch call per step ---------------
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
        h_scores, h_best, h_mapq, h_ntop, h_pfail = pin(max(npairs, 1), torch.float32), pin(n, torch.int32), pin(n, torch.int32), pin(n, torch.int32), pin(n, torch.int32)
        h_recs, h_strings = pin((n, 32), torch.uint8), pin(str_cap, torch.uint8)
        host_in = batch_in(h_packed.data_ptr(), h_len.data_ptr(), exc_ptr, h_cb.data_ptr(), h_desc.data_ptr())
        host_out = BatchOut(h_scores.data_ptr(), h_best.data_ptr(), h_mapq.data_ptr(), h_ntop.data_ptr(), h_pfail.data_ptr() if paired else None, h_recs.data_ptr(),
                            h_strings.data_ptr(), str_cap, 0, None)

        def e2e_step():
            check(lib.ngm_b200_run_batch(ctx, C.byref(host_in), C.byref(host_out)))

        for _ in range(2):
            if paired:
                sw.pe_configure()
            e2e_step()
        barrier()
        e2e_launch0 = sw.launch_count()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        e2e_s = time.perf_counter() - t0
        e2e_ms = sharding.max_over_ranks(e2e_s * 1e3, dev)
        used_e2e = int(host_out.str_used)
        h2d = n * row_bytes + n * 2 + (n + 1) * 4 + npairs * 8 + len(exc_np) * 8
        d2h = npairs * 4 + n * (4 + 4 + 4 + 32) + (n * 4 if paired else 0) + used_e2e + 4 * ((n + args.sub_batch - 1) // args.sub_batch)
        e2e = {"value": world * n * args.steps / (e2e_ms / 1e3), "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms / args.steps, "sub_batch_reads": args.sub_batch, "lanes": args.lanes,
               "h2d_gbs_per_gpu": h2d / (e2e_ms / args.steps * 1e-3) / 1e9, "d2h_gbs_per_gpu": d2h / (e2e_ms / args.steps * 1e-3) / 1e9,
               "api": "one ngm_b200_run_batch call per step: pinned host reads (2 bit/base) + 64-bit descriptors + candidate offsets in; scores, winners, MAPQ, "
                      "alignment records and CIGAR/MD strings back in pinned host memory; the library cuts the batch into sub-batches over its own lanes/streams",
               "gpu_launches": sw.launch_count() - e2e_launch0, "numa": numa,
               "timer": "host wall clock around K calls (barrier + synchronize both sides), max over ranks"}
        if paired:                                   # compare a first batch after configure with the resident first batch
            sw.pe_configure()
            e2e_step()
        h_view = h_recs.numpy().view(ALIGN_REC).reshape(-1)
        fields = ["position_offset", "qstart", "qend", "nm", "identity", "score", "cigar_len", "md_len"]
        same = all(np.array_equal(h_view[f], recs_all[f]) for f in fields) and np.array_equal(h_best.numpy(), best_all) and np.array_equal(h_mapq.numpy(), mapq_all)
        same = same and np.array_equal(h_scores.numpy()[:npairs], d_scores.cpu().numpy()[:npairs])
        # strings: the e2e heap is sparse (one slot per sub-batch), the resident heap dense; compare the text of a sample of reads
        hs = h_strings.numpy()
        for r in range(0, n, max(1, n // 50_000)):
            a, b = h_view[r], recs_all[r]
            if b["score"] >= 0:
                la = int(a["cigar_len"]) + int(a["md_len"])
                same = same and bytes(hs[int(a["str_off"]): int(a["str_off"]) + la]) == bytes(heap_res[int(b["str_off"]): int(b["str_off"]) + la])
        e2e["matches_resident"] = bool(same)

    # ---- candidate search on device: reads -> k-mer vote -> candidates -> score -> top1 -> align ----
    cs_info = None
    if not args.no_cs:
        try:
            from nextgenmap_b200.host.cuda_sw import CsParams, _CContigRec
            import ctypes as C
            K_MER, K_SKIP, BIN = 13, 2, 2
            contig_arr = (_CContigRec * args.contigs)()
            for i, s0 in enumerate(ref.contig_start):
                contig_arr[i].start, contig_arr[i].length, contig_arr[i].name_len = int(s0), int(ref.contig_len), 0
            csp = CsParams(K_MER, K_SKIP, BIN, 1, args.sensitivity, 0.0, 0, 0)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if rank == 0 or not distributed:
                check(lib.ngm_b200_cs_build_index(ctx, C.byref(csp), contig_arr, args.contigs))
            torch.cuda.synchronize()
            build_s = time.perf_counter() - t0
            t0 = time.perf_counter()
            info = sharding.broadcast_index_device(sw, csp, dev, src=0)      # rank 0's prefix table goes to every GPU over NCCL
            torch.cuda.synchronize()
            index_bcast_s = time.perf_counter() - t0 if distributed else 0.0
            cap = 3 * n + 1024
            d_cb = torch.empty(n + 1, dtype=torch.int32, device=dev)
            d_cpairs = torch.empty((cap, 16), dtype=torch.uint8, device=dev)
            d_cvotes = torch.empty(cap, dtype=torch.float32, device=dev)
            d_cscores = torch.empty(cap, dtype=torch.float32, device=dev)

            def cs_only():
                check(lib.ngm_b200_dev_cs_search(ctx, batch.reads.data_ptr(), n, qml, 0, d_cb.data_ptr(), d_cpairs.data_ptr(), d_cvotes.data_ptr(), cap, None, st))

            def cs_pipeline_step():
                check(lib.ngm_b200_dev_set_reads(ctx, batch.reads.data_ptr(), n, qml, st))
                cs_only()
                total = int(d_cb[n].item())                      # the one host sync of the step: number of candidates
                if total > cap:
                    raise RuntimeError(f"candidate buffer too small ({total} > {cap})")
                check(lib.ngm_b200_dev_score_pairs(ctx, MODE_LOCAL, total, d_cpairs.data_ptr(), d_cscores.data_ptr(), st))
                check(lib.ngm_b200_dev_select_top1(ctx, n, d_cb.data_ptr(), d_cscores.data_ptr(), d_best.data_ptr(), d_mapq.data_ptr(), st))
                check(lib.ngm_b200_dev_gather_winners_scored(ctx, n, d_cpairs.data_ptr(), d_cscores.data_ptr(), d_best.data_ptr(), d_wpairs.data_ptr(),
                                                             d_wscores.data_ptr(), st))
                d_cursor.zero_()
                check(lib.ngm_b200_dev_align_pairs_scored(ctx, MODE_LOCAL, n, d_wpairs.data_ptr(), d_wscores.data_ptr(), d_recs.data_ptr(),
                                                          d_strings.data_ptr(), str_cap, d_cursor.data_ptr(), st))
                return total

            cs_only()                                               # size the candidate buffers for this workload (divergent reads: ~20 per read)
            need = int(d_cb[n].item())
            if need > cap:
                cap = int(need * 1.05) + 1024
                del d_cpairs, d_cvotes, d_cscores
                d_cpairs = torch.empty((cap, 16), dtype=torch.uint8, device=dev)
                d_cvotes = torch.empty(cap, dtype=torch.float32, device=dev)
                d_cscores = torch.empty(cap, dtype=torch.float32, device=dev)
            ms_cs = time_call(cs_only, reps=3)
            total_c = cs_pipeline_step()
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a_.record()
            reps = max(2, min(args.steps, 5))
            for _ in range(reps):
                cs_pipeline_step()
            b_.record()
            barrier()
            ms_pipe = sharding.max_over_ranks(a_.elapsed_time(b_) / reps, dev)
            # sanity against the generator's truth: the winner of every read sits at the locus the read was drawn from
            wp = d_wpairs.view(torch.int64).view(n, 2)
            loc = wp[:, 0] + (corridor >> 1)
            wflags = (wp[:, 1] >> 32) & 0xFFFFFFFF
            ok_pos = ((loc - batch.true_pos).abs() <= corridor) & ((wflags & 1).bool() == batch.reverse) & ((wflags & 4) == 0)
            recs_cs = d_recs.cpu().numpy().view(ALIGN_REC).reshape(-1)
            n_kmers = L - K_MER + 1
            mean_list = info["table_len"] / float(4 ** K_MER)
            alg_bytes_read = n_kmers * 2 * 8 + n_kmers * 2 * mean_list * 4 + L
            cs_info = {"kmer": K_MER, "kmer_skip": K_SKIP, "bin_size": BIN, "sensitivity": args.sensitivity, "max_kfreq": info["max_kfreq"],
                       "index_positions": info["table_len"], "index_build_seconds": build_s, "index_broadcast_seconds": index_bcast_s, "index_bytes": 4 * (4 ** K_MER + 1) + 4 * info["table_len"],
                       "cs_ms": ms_cs, "cs_reads_per_s": n / (ms_cs * 1e-3), "candidates": int(total_c), "candidates_per_read": total_c / n,
                       "pipeline_step": "set_reads -> cs_search (k-mer vote) -> score all candidates -> top1+MAPQ -> gather -> align+backtrace+CIGAR/MD",
                       "pipeline_ms": ms_pipe, "pipeline_reads_per_s": world * n / (ms_pipe * 1e-3),
                       "winner_at_true_locus": float(ok_pos.float().mean().item()), "mapped": int(np.count_nonzero(recs_cs["score"] >= 0)),
                       "reads_on_sequential_exact_kernel": sw.cs_exact_reads(), "exact_kernel_reasons": sw.cs_exact_reasons(),
                       "roofline": {"kernel": "cs_search_kernel", "bound": "hbm", "algorithmic_bytes_per_read": alg_bytes_read,
                                    "achieved": alg_bytes_read * n / (ms_cs * 1e-3) / 1e9, "unit": "GB/s",
                                    "note": "one 16-byte index entry and 2 position lists of ~%.1f x 4 B per k-mer, %d k-mers per read" % (mean_list, n_kmers)}}
            cs_info["roofline"]["peak"] = hbm_peak_cs = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6542.0)) if (ROOT / "MEASURED_PEAKS.json").exists() else 6542.0
            cs_info["roofline"]["frac"] = cs_info["roofline"]["achieved"] / hbm_peak_cs
            tfc = ROOT / "profiles" / "traffic_r1.json"
            if tfc.exists() and L == READ_LEN:
                cs_info["roofline"]["traffic"] = json.loads(tfc.read_text())["cs_search_kernel"]["dram_bytes_per_unit"] * n
            # the whole mapping step through the one-call C ABI entry point, host buffers in and out (ngm_b200_map_batch: sub-batches over the
            # library's lanes, upload / candidate search / score + align / download overlapped), rank 0
            if rank == 0 and not args.no_e2e:
                try:
                    from nextgenmap_b200.host.cuda_sw import MapResult
                    h_all = torch.empty((n, qml), dtype=torch.uint8).pin_memory()
                    h_all.copy_(batch.reads)
                    capm, scapm = int(total_c * 1.02) + 4096, n * STR_PER + 4096
                    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
                    hb = {"begin": pin(n + 1, torch.int32), "pairs": pin((capm, 16), torch.uint8), "scores": pin(capm, torch.float32), "best": pin(n, torch.int32),
                          "mapq": pin(n, torch.int32), "ntop": pin(n, torch.int32), "mh": pin(n, torch.float32), "recs": pin((n, 32), torch.uint8),
                          "heap": pin(scapm, torch.uint8)}
                    resm = MapResult(hb["begin"].data_ptr(), hb["pairs"].data_ptr(), hb["scores"].data_ptr(), capm, 0, hb["best"].data_ptr(), hb["mapq"].data_ptr(),
                                     hb["ntop"].data_ptr(), None, hb["mh"].data_ptr(), hb["recs"].data_ptr(), hb["heap"].data_ptr(), scapm, 0)

                    def map_all():
                        check(lib.ngm_b200_map_batch(ctx, h_all.data_ptr(), n, qml, MODE_LOCAL, 0, C.byref(resm)))
                        return 4 * (n + 1) + int(resm.n_candidates) * 20 + n * (16 + 32) + int(resm.str_used)
                    map_all()
                    t0 = time.perf_counter()
                    d2h_m = map_all()
                    map_s = time.perf_counter() - t0
                    same_m = bool(np.array_equal(hb["best"].numpy(), d_best.cpu().numpy()) and np.array_equal(hb["recs"].numpy().view(ALIGN_REC).reshape(-1)["position_offset"],
                                                                                                             recs_cs["position_offset"]))
                    cs_info["e2e_map_batch"] = {"value": n / map_s, "unit": "reads/s", "reads": n, "sub_batch_reads": args.sub_batch, "lanes": args.lanes, "h2d_bytes": n * qml,
                                                "d2h_bytes": d2h_m, "matches_resident_pipeline": same_m,
                                                "note": "ONE ngm_b200_map_batch call: pinned host ASCII reads -> candidates, scores, selection, alignments, CIGAR/MD back on the host; "
                                                        "the library pipelines sub-batches over its lanes"}
                    del h_all, hb
                except Exception as e:  # noqa: BLE001
                    cs_info["e2e_map_batch"] = {"error": str(e)}
            # SAM records of the single-end run on the host threads (ngm_b200_format_sam): 1 M reads of the step above
            if rank == 0:
                try:
                    from nextgenmap_b200.host.cuda_sw import SamBatch, SamOpts, _CContig, _CEncRef
                    n_f = min(n, 1_000_000)
                    e_f = int(d_cb[n_f].item())
                    h_reads = batch.reads[:n_f].cpu().numpy()
                    h_quals = np.where(h_reads != 0, ord("I"), 0).astype(np.uint8)
                    name_arr = (C.c_char_p * n_f)(*[b"r%d" % i for i in range(n_f)])
                    hk = [d_cpairs[:e_f].cpu().numpy(), d_cscores[:e_f].cpu().numpy(), d_best[:n_f].cpu().numpy(), d_mapq[:n_f].cpu().numpy(),
                          np.ones(n_f, np.int32), np.zeros(n_f, np.float32), d_recs.cpu().numpy().view(np.uint8).reshape(n, -1)[:n_f].copy(), d_strings.cpu().numpy()]
                    ctg = (_CContig * args.contigs)()
                    for i, s0 in enumerate(ref.contig_start):
                        ctg[i].start, ctg[i].length, ctg[i].name_len, ctg[i].name = int(s0), int(ref.contig_len), len("chr%d" % (i + 1)), b"chr%d" % (i + 1)
                    enc = _CEncRef()
                    enc.concat_len, enc.packed_bytes, enc.n_contigs = ref.concat_len, 0, args.contigs
                    enc.contigs = C.cast(ctg, type(enc.contigs))
                    sb = SamBatch(n_f, qml, h_reads.ctypes.data, h_quals.ctypes.data, name_arr, hk[0].ctypes.data, hk[1].ctypes.data, hk[2].ctypes.data, hk[3].ctypes.data,
                                  hk[4].ctypes.data, None, hk[5].ctypes.data, hk[6].ctypes.data, hk[7].ctypes.data)
                    so = SamOpts(0.65, 0.5, 0, 1000, 0)
                    out_f = np.zeros(n_f * (2 * qml + 256), np.uint8)
                    used = C.c_size_t(0)
                    fmt_t = []
                    for _ in range(4):                         # a writer formats batch after batch into the same buffer: the first call (which
                        t0 = time.perf_counter()               # first-touches the output buffer and sizes the library's parts) is reported apart
                        rc_f = lib.ngm_b200_format_sam(C.byref(enc), C.byref(so), C.byref(sb), out_f.ctypes.data, out_f.size, C.byref(used))
                        fmt_t.append(time.perf_counter() - t0)
                    fmt_s = float(np.median(fmt_t[1:]))
                    cs_info["sam_format"] = {"reads": n_f, "host_threads": host_threads, "reads_per_s": n_f / fmt_s, "first_call_reads_per_s": n_f / fmt_t[0],
                                             "bytes": int(used.value), "rc": int(rc_f),
                                             "mapped_lines": int(out_f[: used.value].tobytes().count(b"\tAS:i:")),
                                             "note": "ngm_b200_format_sam on the host: AlignmentBuffer::WriteRead + GenericReadWriter filters + SAMWriter lines; "
                                                     "median of 3 calls after the first"}
                    del out_f, hk, h_reads, h_quals
                except Exception as e:  # noqa: BLE001
                    cs_info["sam_format"] = {"error": str(e)}
            # parity at full scale + CPU beside it (rank 0): the oracle restatement of CS.cpp searches a sample of the reads in
            # the SAME 3 Gbp prefix table (exported from the device); lists must agree entry by entry, order included
            if rank == 0 and not args.no_cpu_baseline:
                try:
                    from oracle import cs_port
                    n_cs = min(n, 20_000)
                    tab_h, weight_h, table_h = sw.cs_export_index()
                    oix = cs_port.Index.from_arrays(tab_h, weight_h, table_h, K_MER, K_SKIP, BIN, info["max_kfreq"])
                    reads_cs = batch.reads[:n_cs].cpu().numpy()
                    t0 = time.perf_counter()
                    wb, wc, wm = oix.search(reads_cs, args.sensitivity)
                    cpu_s = time.perf_counter() - t0
                    gb = d_cb[: n_cs + 1].cpu().numpy()
                    gp = d_cpairs[: int(gb[-1])].cpu().numpy().view(np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])).reshape(-1)
                    gv = d_cvotes[: int(gb[-1])].cpu().numpy()
                    same = (np.array_equal(gb, wb) and np.array_equal((gp["window_start"] + np.uint64(corridor >> 1)).astype(np.uint64), wc["location"])
                            and np.array_equal(gp["flags"] & 1, wc["reverse"].astype(np.uint32)) and np.array_equal(gv, wc["score"]))
                    cs_info["parity_sample"] = {"reads_checked": n_cs, "candidates_checked": int(wb[-1]), "identical_to_oracle": bool(same)}
                    cs_info["cpu_baseline"] = {"value": n_cs / cpu_s, "unit": "reads/s", "cores": 1, "kind": "port",
                                               "sample": f"{n_cs} reads of the same workload searched by oracle/cs_oracle.c (restatement of CS.cpp) in the same prefix table"}
                    del tab_h, weight_h, table_h, oix
                except Exception as e:  # noqa: BLE001
                    cs_info["parity_sample"] = {"error": str(e)}
            # the reference's whole program on the host cores beside the device pipeline, rank 0, single GPU runs only
            if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.no_ngm and L == READ_LEN and not args.corridor:
                try:
                    cs_info["cpu_baseline_ngm"] = run_ngm_whole(sw, ref, batch, min(n, args.ngm_sample), L, qml, args.contigs, args.sensitivity, host_threads, lib)
                except Exception as e:  # noqa: BLE001
                    cs_info["cpu_baseline_ngm"] = {"error": str(e)[-400:]}
            # ---- paired-end (BASELINE configs[2] shape, per-GPU shard): the same run with mates in rows 2f / 2f + 1 and
            # ScoreBuffer::top1PE on the device (ngm_b200_dev_select_pairs) between scoring and alignment
            if not args.no_pe:
                try:
                    from nextgenmap_b200.host.cuda_sw import PeParams
                    pbatch = workload.make_reads(ref, n, L, qml, corridor, seed=20261018 + 2 + 16 * rank, sub_rate=args.sub_rate, indel_rate=args.indel_rate,
                                                 paired=True)
                    d_nt, d_pf = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.int32, device=dev)
                    pep = PeParams(0.9, 0, 1000, 0, 0)

                    def pe_select(total):
                        check(lib.ngm_b200_dev_select_pairs(ctx, n, d_cb.data_ptr(), d_cpairs.data_ptr(), d_cscores.data_ptr(), total, d_best.data_ptr(),
                                                            d_mapq.data_ptr(), d_nt.data_ptr(), d_pf.data_ptr(), st))

                    def pe_pipeline_step():
                        check(lib.ngm_b200_dev_set_reads(ctx, pbatch.reads.data_ptr(), n, qml, st))
                        check(lib.ngm_b200_dev_cs_search(ctx, pbatch.reads.data_ptr(), n, qml, 0, d_cb.data_ptr(), d_cpairs.data_ptr(), d_cvotes.data_ptr(), cap, None, st))
                        total = int(d_cb[n].item())
                        if total > cap:
                            raise RuntimeError(f"candidate buffer too small ({total} > {cap})")
                        check(lib.ngm_b200_dev_score_pairs(ctx, MODE_LOCAL, total, d_cpairs.data_ptr(), d_cscores.data_ptr(), st))
                        pe_select(total)
                        check(lib.ngm_b200_dev_gather_winners_scored(ctx, n, d_cpairs.data_ptr(), d_cscores.data_ptr(), d_best.data_ptr(), d_wpairs.data_ptr(),
                                                                     d_wscores.data_ptr(), st))
                        d_cursor.zero_()
                        check(lib.ngm_b200_dev_align_pairs_scored(ctx, MODE_LOCAL, n, d_wpairs.data_ptr(), d_wscores.data_ptr(), d_recs.data_ptr(),
                                                                  d_strings.data_ptr(), str_cap, d_cursor.data_ptr(), st))
                        return total

                    check(lib.ngm_b200_pe_configure(ctx, C.byref(pep)))
                    total_p = pe_pipeline_step()                         # first batch after configure: checked against the oracle below
                    torch.cuda.synchronize()
                    first = {"begin": d_cb.cpu().numpy(), "best": d_best.cpu().numpy(), "mapq": d_mapq.cpu().numpy(), "nt": d_nt.cpu().numpy(), "pf": d_pf.cpu().numpy()}
                    sum1, cnt1 = sw.pe_insert_stats()
                    deferred1 = sw.pe_deferred_fragments()
                    wp = d_wpairs.view(torch.int64).view(n, 2)
                    wflags = (wp[:, 1] >> 32) & 0xFFFFFFFF
                    ok_pe = (((wp[:, 0] + (corridor >> 1)) - pbatch.true_pos).abs() <= corridor) & ((wflags & 1).bool() == pbatch.reverse) & ((wflags & 4) == 0)
                    ms_sel = time_call(lambda: pe_select(total_p), reps=3)
                    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    barrier()
                    a_.record()
                    for _ in range(reps):
                        pe_pipeline_step()
                    b_.record()
                    barrier()
                    ms_pe = sharding.max_over_ranks(a_.elapsed_time(b_) / reps, dev)
                    pe_info = {"workload": f"{n // 2} fragments x 2 x {L} bp per GPU, insert ~ N(400, 40), FR, same reference and index",
                               "pipeline_step": "set_reads -> cs_search -> score all candidates -> top1PE (select_pairs) -> gather -> align+backtrace+CIGAR/MD",
                               "candidates_per_read": total_p / n, "select_pairs_ms": ms_sel, "pipeline_ms": ms_pe, "pipeline_reads_per_s": world * n / (ms_pe * 1e-3),
                               "proper_pair_fraction": float(((first["pf"] == 0) & (first["best"] >= 0)).mean()), "winner_at_true_locus": float(ok_pe.float().mean().item()),
                               "mean_insert_size": sum1 / max(cnt1, 1), "pairs_accepted": cnt1 - 1,
                               "fragments_needing_insert_mean": deferred1}
                    if rank == 0 and not args.no_cpu_baseline:
                        try:
                            from oracle import mapper_port
                            n_pe = min(n, 20_000)
                            e_pe = int(first["begin"][n_pe])
                            pdt = np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])
                            # candidates and scores of the first batch: recompute them (the buffers were overwritten by the timed steps)
                            check(lib.ngm_b200_dev_set_reads(ctx, pbatch.reads.data_ptr(), n, qml, st))
                            check(lib.ngm_b200_dev_cs_search(ctx, pbatch.reads.data_ptr(), n, qml, 0, d_cb.data_ptr(), d_cpairs.data_ptr(), d_cvotes.data_ptr(), cap, None, st))
                            check(lib.ngm_b200_dev_score_pairs(ctx, MODE_LOCAL, total_p, d_cpairs.data_ptr(), d_cscores.data_ptr(), st))
                            torch.cuda.synchronize()
                            hp = d_cpairs[:e_pe].cpu().numpy().view(pdt).reshape(-1)
                            hs = d_cscores[:e_pe].cpu().numpy()
                            lens = np.full(n_pe, L, np.int32)
                            t0 = time.perf_counter()
                            want = mapper_port.Selector().select_pairs(first["begin"][: n_pe + 1], hp["window_start"] + np.uint64(corridor >> 1), hs, lens)
                            cpu_s = time.perf_counter() - t0
                            has = want["best"] >= 0
                            same = (np.array_equal(first["best"][:n_pe], want["best"]) and np.array_equal(first["mapq"][:n_pe], want["mapq"])
                                    and np.array_equal(first["nt"][:n_pe][has], want["num_top"][has]) and np.array_equal(first["pf"][:n_pe], want["paired_fail"]))
                            pe_info["parity_sample"] = {"reads_checked": n_pe, "identical_to_oracle": bool(same)}
                            pe_info["cpu_baseline"] = {"value": n_pe / cpu_s, "unit": "reads/s", "cores": 1, "kind": "port",
                                                       "sample": f"selection only (oracle/select_oracle.c) of {n_pe} reads with the device's scores"}
                        except Exception as e:  # noqa: BLE001
                            pe_info["parity_sample"] = {"error": str(e)}
                    cs_info["paired_end"] = pe_info
                    del pbatch
                except Exception as e:  # noqa: BLE001
                    cs_info["paired_end"] = {"error": str(e)}
            del d_cpairs, d_cvotes, d_cscores, d_cb
        except Exception as e:  # noqa: BLE001
            cs_info = {"error": str(e)}

    # ---- reductions over ranks ----------------------------------------------
    ms_max = sharding.max_over_ranks(ms, dev)
    ctr = sharding.reduce_counters({"reads": n, "mapped": mapped, "pairs_scored": npairs}, dev)      # NGM.cpp:172-201
    total_reads = ctr["reads"]

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- parity spot check + CPU baseline (rank 0, outside the timed regions) ----
    parity = None
    cpu_baseline = None
    try:
        from oracle import port
        n_s = args.cpu_sample or min(n, max(20_000, 3000 * host_threads))
        packed_host = ref.packed.cpu().numpy()
        reads_h = batch.reads[:n_s].cpu().numpy()
        cb = batch.cand_begin[: n_s + 1].cpu().numpy()
        pdt = np.dtype([("window_start", "<u8"), ("read_index", "<u4"), ("flags", "<u4")])
        pairs_h = batch.pairs[: int(cb[-1])].cpu().numpy().view(pdt).reshape(-1)
        refs_s, qrys_s, n_align = cpu_sample_windows(packed_host, ref.concat_len, reads_h, pairs_h, cb, n_s, qml, corridor)
        # order used by cpu_sample_windows: first candidates of every read, then the rest
        first = [int(cb[r]) for r in range(n_s)]
        rest = [j for r in range(n_s) for j in range(int(cb[r]) + 1, int(cb[r + 1]))]
        order = np.array(first + rest)
        want = port.batch_score(refs_s, qrys_s, qml, corridor, MODE_LOCAL)
        got = d_scores[: int(cb[-1])].cpu().numpy()[order]
        parity = {"score_pairs_checked": int(len(order)), "score_mismatches": int(np.count_nonzero(want != got))}
        # alignments of the sample's winners: position, QStart / QEnd, NM, identity, CIGAR and MD against the oracle's BatchAlign
        # (windows the way AlignmentBuffer decodes them: refMaxLen of AlignmentBuffer.h:67)
        inv = np.empty(len(order), np.int64)
        inv[order] = np.arange(len(order))
        win = best_all[:n_s]
        sel_reads = np.nonzero(win >= 0)[0][:25_000]
        abuf = (qml + corridor) | 2
        refs_a = np.zeros((len(sel_reads), max(abuf, qml + corridor)), np.uint8)
        for i, r in enumerate(sel_reads):
            w = port.decode_window(packed_host, ref.concat_len, int(pairs_h["window_start"][win[r]]), abuf)
            if w is not None:
                refs_a[i, :abuf] = np.frombuffer(w, np.uint8)
        qrys_a = qrys_s[inv[win[sel_reads]]]
        wa = port.batch_align(refs_a, qrys_a, qml, corridor, MODE_LOCAL)
        bad = 0
        for i, r in enumerate(sel_reads):
            g, a = recs_all[r], wa[i]
            o = int(g["str_off"])
            cig = bytes(heap_res[o: o + int(g["cigar_len"])])
            md = bytes(heap_res[o + int(g["cigar_len"]): o + int(g["cigar_len"]) + int(g["md_len"])]).split(b"\0")[0]
            if a.ascore == -1.0 and a.cigar == b"!!!":
                ok = float(g["score"]) == -1.0
            else:
                ok = ((int(g["position_offset"]), int(g["qstart"]), int(g["qend"]), int(g["nm"]), np.float32(g["identity"]).tobytes(), float(g["score"]), cig, md)
                      == (a.position_offset, a.qstart, a.qend, a.nm, np.float32(a.identity).tobytes(), a.ascore, a.cigar, a.md))
            bad += 0 if ok else 1
        parity.update({"alignments_checked": int(len(sel_reads)), "align_mismatches": int(bad),
                       "align_fields": "PositionOffset, QStart, QEnd, NM, Identity (bit pattern), Align.Score, CIGAR, MD vs oracle/ngm_oracle.c BatchAlign"})
        if not args.no_cpu_baseline:
            res = run_cpu_reference(refs_s, qrys_s, n_align, n_s, qml, corridor, host_threads, 1, 3)
            cpu_baseline = {"value": res["value"], "unit": "reads/s", "cores": res["cores"], "kind": res["kind"],
                            "sample": f"{n_s} reads / {len(refs_s)} pairs of the same workload, 3 timed passes after 1 warm pass",
                            "score_pairs_per_s": res.get("score_pairs_per_s"), "align_pairs_per_s": res.get("align_pairs_per_s")}
    except Exception as e:  # noqa: BLE001
        import traceback
        parity = {"error": str(e), "trace": traceback.format_exc()[-600:]}

    # ---- strict IAlignment path (char** in, struct Align out), one host thread, for the record ----
    strict = None
    try:
        n_sp = min(262144, npairs)
        pv = batch.pairs[:n_sp]
        starts = pv[:, 0:8].contiguous().view(torch.int64).reshape(-1)
        ridx = pv[:, 8:12].contiguous().view(torch.int32).reshape(-1).long()
        rev = (pv[:, 12:16].contiguous().view(torch.int32).reshape(-1) & 1).bool()
        width = ((qml + corridor) | 1) + 1
        pos = starts[:, None] + torch.arange(width - 2, device=dev)[None, :]
        b = ref.packed[pos >> 1]
        codes = torch.where((pos & 1) == 1, b & 0xF, b >> 4).long()
        lut = workload.ASCII_OF_NGM.to(dev)
        win = torch.zeros((n_sp, width), dtype=torch.uint8, device=dev)
        win[:, : width - 2] = lut[codes]
        rd = batch.reads[ridx]
        comp = torch.arange(256, dtype=torch.uint8, device=dev)
        for a_, b_ in zip(b"ACGT", b"TGCA"):
            comp[a_] = b_
        rc = torch.zeros_like(rd)
        rc[:, :L] = comp[torch.flip(rd[:, :L], dims=[1]).long()]
        rd = torch.where(rev[:, None], rc, rd)
        refs_h, qrys_h = win.cpu().numpy(), rd.cpu().numpy()
        ssw = CudaSW(qml, corridor, device=local_rank)
        ssw.BatchScore(MODE_LOCAL, refs_h[:4096], qrys_h[:4096])
        t0 = time.perf_counter()
        sc_strict = ssw.BatchScore(MODE_LOCAL, refs_h, qrys_h)
        t1 = time.perf_counter()
        n_al = n_sp // 2
        # the caller's result buffers exist before the call, as in AlignmentBuffer (allocated once per thread, AlignmentBuffer.cpp:106-109):
        # the timed region is the C call, not the harness's allocation and first touch of ~190 MB of CIGAR / MD rows
        bufs = ssw.alloc_align_buffers(n_al)
        rows = lambda a: a.ctypes.data + np.arange(a.shape[0], dtype=np.uint64) * np.uint64(a.strides[0])
        rp_al, qp_al = rows(refs_h[:n_al]), rows(qrys_h[:n_al])
        ssw.batch_align_raw(MODE_LOCAL, refs_h[:4096], qrys_h[:4096], buffers=bufs)
        ssw.batch_align_raw(MODE_LOCAL, refs_h[:n_al], qrys_h[:n_al], buffers=bufs)          # (warm: staging buffers sized, result rows touched)
        al_s = []
        for _ in range(3):
            t2 = time.perf_counter()
            got_al = ssw.lib.ngm_b200_batch_align(ssw.ctx, MODE_LOCAL, n_al, rp_al.ctypes.data, qp_al.ctypes.data, qp_al.ctypes.data, bufs[0].ctypes.data, None)
            al_s.append(time.perf_counter() - t2)
            assert got_al == n_al
        sc_s = []
        rp_sc, qp_sc = rows(refs_h), rows(qrys_h)
        out_sc = np.zeros(n_sp, np.float32)
        for _ in range(3):
            t2 = time.perf_counter()
            got_sc = ssw.lib.ngm_b200_batch_score(ssw.ctx, MODE_LOCAL, n_sp, rp_sc.ctypes.data, qp_sc.ctypes.data, out_sc.ctypes.data, None)
            sc_s.append(time.perf_counter() - t2)
            assert got_sc == n_sp
        same = bool(np.array_equal(sc_strict, d_scores[:n_sp].cpu().numpy())) and bool(np.array_equal(out_sc, sc_strict))
        strict = {"score_pairs_per_s": n_sp / float(np.median(sc_s)), "align_pairs_per_s": n_al / float(np.median(al_s)), "host_threads": 1,
                  "pairs": n_sp, "align_pairs": n_al, "scores_equal_descriptor_path": same,
                  "first_call_score_pairs_per_s": n_sp / (t1 - t0),
                  "note": "drop-in IAlignment::BatchScore/BatchAlign with host char** buffers, one host thread, one call each (median of 3 after a warm "
                          "call; the caller's result buffers are allocated before the timed region, as AlignmentBuffer does); host gather + PCIe bound"}
        ssw.close()
        # NGM drives the backend from all of its CS threads at once, one IAlignment instance each (CS.cpp:455-461): the same calls from T host
        # threads, every thread with its own context and its own slice of the pairs (ctypes releases the GIL inside the calls)
        import threading
        T = max(1, min(8, host_threads))
        per_sc, per_al = (n_sp // T) & ~3, (n_al // T) & ~3
        if T > 1 and per_al >= 4096:
            ctxs = [CudaSW(qml, corridor, device=local_rank) for _ in range(T)]
            tb = [c_.alloc_align_buffers(per_al) for c_ in ctxs]
            outs = [np.zeros(per_sc, np.float32) for _ in range(T)]
            ptrs = [(rows(refs_h[t * per_sc:(t + 1) * per_sc]), rows(qrys_h[t * per_sc:(t + 1) * per_sc])) for t in range(T)]

            def work(t, what):
                c_, (rp_, qp_) = ctxs[t], ptrs[t]
                if what == "score":
                    assert c_.lib.ngm_b200_batch_score(c_.ctx, MODE_LOCAL, per_sc, rp_.ctypes.data, qp_.ctypes.data, outs[t].ctypes.data, None) == per_sc
                else:
                    assert c_.lib.ngm_b200_batch_align(c_.ctx, MODE_LOCAL, per_al, rp_.ctypes.data, qp_.ctypes.data, qp_.ctypes.data, tb[t][0].ctypes.data, None) == per_al

            def timed(what):
                th = [threading.Thread(target=work, args=(t, what)) for t in range(T)]
                t_0 = time.perf_counter()
                for x in th:
                    x.start()
                for x in th:
                    x.join()
                return time.perf_counter() - t_0
            for what in ("score", "align"):
                timed(what)                                    # warm: staging buffers of every context
            mt_sc = float(np.median([timed("score") for _ in range(3)]))
            mt_al = float(np.median([timed("align") for _ in range(3)]))
            strict["threads"] = {"host_threads": T, "score_pairs_per_s": T * per_sc / mt_sc, "align_pairs_per_s": T * per_al / mt_al,
                                 "scores_equal": bool(all(np.array_equal(outs[t], sc_strict[t * per_sc:(t + 1) * per_sc]) for t in range(T))),
                                 "note": "the same strict calls from T host threads at once, one context per thread (how NGM's CS threads use the backend)"}
            for c_ in ctxs:
                c_.close()
    except Exception as e:  # noqa: BLE001
        strict = {"error": str(e)} if strict is None else dict(strict, threads_error=str(e))

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    if cs_info and "roofline" in cs_info:
        cs_info["roofline"]["peak"] = hbm_peak
        cs_info["roofline"]["frac"] = cs_info["roofline"]["achieved"] / hbm_peak
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = clk.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    # The step's DP work: the score kernel runs on the candidates of multi-candidate reads only (single-end), the align launch set on every read.
    multi_pairs = npairs if paired else int((batch.cand_begin[1:] - batch.cand_begin[:-1])[(batch.cand_begin[1:] - batch.cand_begin[:-1]) > 1].sum().item())
    ms_score_step = ms_score * multi_pairs / max(npairs, 1)
    dominant = "score" if ms_score_step >= ms_align else "align"
    dom_ms = ms_score if dominant == "score" else ms_align
    dom_units = npairs if dominant == "score" else n
    alg_bytes = ALG_BYTES_SCORE * dom_units if dominant == "score" else (ALG_BYTES_SCORE + 8 + 2 * 4) * dom_units
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM traffic of the dominant launch set from the committed ncu capture (bytes per unit x units of this run)
    traffic, traffic_src = None, None
    for name in ("traffic_r2.json", "traffic_r1.json"):
        tf = ROOT / "profiles" / name
        if tf.exists() and L == READ_LEN and not args.corridor:
            tj = json.loads(tf.read_text())
            fwd_key = "align_s16_fwd2_kernel" if "align_s16_fwd2_kernel" in tj else "align_s16_fwd_kernel"
            per_unit = tj["score_s16_kernel"]["dram_bytes_per_unit"] if dominant == "score" else (
                tj[fwd_key]["dram_bytes_per_unit"] + tj["backtrace_format_kernel"]["dram_bytes_per_unit"])
            traffic, traffic_src = per_unit * dom_units, f"profiles/{name}"
            break
    # Integer-ALU roofline against the MEASURED issue rate of the recurrence's own instruction (VIADDMNMX.S16x2, ngm_b200_alu_peak):
    # one s16x2 instruction serves two cells.  Floors: score 4 ALU instructions per cell pair (substitution PRMT, diag VIADD, two VIADDMNMX);
    # forward-with-pointers 5 (+ the LOP3 that strips the direction tag; the tag arithmetic itself runs on the FMA pipe).
    score_gcups = npairs * CELLS_PER_PAIR / (ms_score * 1e-3) / 1e9
    align_gcups = n * CELLS_PER_PAIR / (ms_align * 1e-3) / 1e9
    fwd_gcups = n * CELLS_PER_PAIR / (max(ms_fwd, 1e-6) * 1e-3) / 1e9
    peak_score_gcups = alu_rate * 2 / 4 / 1e9
    peak_fwd_gcups = alu_rate * 2 / 5 / 1e9
    survey_peak_gcups = 2 * sm_count * 128 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 5 / 1e9
    line = {
        "metric": "reads/sec aligned (150bp SE vs 3Gbp ref)", "value": total_reads * args.steps / (ms_max * 1e-3), "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic", "config": cfg,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        "roofline": {"kernel": "score_s16_kernel" if dominant == "score" else "align_s16_fwd2_kernel + backtrace_format_kernel (one launch set per 1 Mi alignments)",
                     "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": traffic, "traffic_note": f"dram bytes for the units of one launch set pass, from the ncu --set full capture summarised in {traffic_src}",
                     "peak_source": peak_src, "ms_per_launch_set": dom_ms, "units_per_launch_set": dom_units,
                     "algorithmic_bytes_per_unit": alg_bytes / dom_units,
                     "note": "integer DP: ~22 cells per algorithmic byte, so the HBM fraction is small by construction; the binding roofline is roofline_alu"},
        "roofline_alu": {"bound": "int-alu issue", "measured_viaddmnmx_s16x2_per_s": alu_rate, "measured_imad_per_s": imad_rate, "measured_mixed_per_s": mixed_rate,
                         "score_gcups": score_gcups, "peak_score_gcups": peak_score_gcups, "score_frac": score_gcups / peak_score_gcups,
                         "align_forward_gcups": fwd_gcups, "peak_forward_gcups": peak_fwd_gcups, "align_forward_frac": fwd_gcups / peak_fwd_gcups,
                         "align_launch_set_gcups": align_gcups, "align_launch_set_frac": align_gcups / peak_fwd_gcups,
                         "step_forward_gcups": (step_fwd_pairs * CELLS_PER_PAIR / (step_fwd_ms * 1e-3) / 1e9) if step_fwd_ms else None,
                         "issue_model": "profiles/r2_issue_rates.md: measured issue rates of every instruction of the inner loop; the forward kernel's loop needs "
                                        "6.1 integer-pipe instructions per slot (two cells) and runs at 86 % of that bound",
                         "survey_model_peak_gcups_s16x2": survey_peak_gcups, "score_frac_of_survey_model": score_gcups / survey_peak_gcups,
                         "align_launch_set_frac_of_survey_model": align_gcups / survey_peak_gcups,
                         "sm_count": sm_count, "sm_mhz_under_load": sm_mhz,
                         "definition": f"cells = L x corridor = {CELLS_PER_PAIR} per pair; peak = measured VIADDMNMX.S16x2 thread-instructions/s x 2 cells / (4 | 5) ALU "
                                       "instructions per cell pair; the survey's model (SMs x 128 lanes x max clock / 5 instr per cell, x2) is kept beside it"},
        "kernel_ms": {"set_reads_packed2": ms_expand, "set_reads_ascii": ms_pack, "score_all_pairs": ms_score, "score_in_step": ms_score_step, "align_launch_sets": ms_align, "align_forward": ms_fwd,
                      "align_backtrace_format": ms_bt, "align_without_known_scores": ms_align_unscored, "launch_sets": launch_sets,
                      "pairs_scored_in_step": multi_pairs, "score_share": ms_score_step / (ms_max / args.steps), "align_share": ms_align / (ms_max / args.steps),
                      "classic_calls_equal_batch": classic_equal, "pack_reads_host_seconds": pack_host_s, "pack_reads_host_reads_per_s": n / pack_host_s,
                      "step_forward_all": step_fwd_ms, "step_pick_backtrace": step_bt_ms, "step_forward_pairs": step_fwd_pairs,
                      "note": "score_* / align_* time the classic entry points on the same data (one kernel each); step_* are the kernels of one resident step"},
        "candidate_search": cs_info,
        "cpu_baseline": cpu_baseline, "parity_sample": parity, "strict_path": strict,
        "counters": {"reads": total_reads, "mapped": ctr["mapped"], "pairs_scored": ctr["pairs_scored"], "string_bytes": used_strings},
        "setup_seconds": setup_s, "host_threads": host_threads,
    }
    emit_json(line)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
