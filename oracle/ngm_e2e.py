"""oracle/ngm_e2e.py -- TEST INFRASTRUCTURE ONLY.

End-to-end runs of the UNMODIFIED NextGenMap built by oracle/Makefile.ngm:
``ngm_ref`` (reference OpenCL backend on the vendored CPU runtime) and ``ngm_cuda`` (same NGM objects, CUDA
backend through the link seam).  BASELINE configs[0]: 10 k x 100 bp single-end synthetic reads vs a 5 Mbp
synthetic reference, ``-t 4``; compare sorted SAM bodies (thread interleaving only permutes lines).
"""
from __future__ import annotations

import os
import subprocess
from pathlib import Path
from typing import List, Sequence

import numpy as np

HERE = Path(__file__).resolve().parent
NGM_DIR = HERE / "_ref" / "ngm"
COMP = bytes.maketrans(b"ACGT", b"TGCA")


def binary(which: str) -> Path:
    return NGM_DIR / f"ngm_{which}"


def available(which: str) -> bool:
    return binary(which).exists()


def write_inputs(d: Path, ref_len: int = 5_000_000, n_reads: int = 10_000, read_len: int = 100, seed: int = 20261017,
                 sub_rate: float = 0.02, indel_reads: float = 0.05, ins_rate: float = 0.0, del_rate: float = 0.0) -> None:
    """Seeded synthetic reference (one contig) + reads: uniform start, both strands, 2 % substitutions,
    a 1-3 bp indel in 5 % of the reads, constant qualities; truth in the read name.
    ins_rate / del_rate: additionally one-base insertions / deletions PER BASE (BASELINE configs[3]: 12 % substitutions + 1.5 % + 1.5 %)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    ref = acgt[rng.integers(0, 4, ref_len)]
    with open(d / "ref.fa", "wb") as f:
        f.write(b">chr1 synthetic\n")
        for i in range(0, ref_len, 60):
            f.write(ref[i: i + 60].tobytes() + b"\n")
    with open(d / "reads.fq", "wb") as f:
        for r in range(n_reads):
            pos = int(rng.integers(0, ref_len - read_len - 8))
            seq = ref[pos: pos + read_len + 4 + (read_len // 8 if del_rate else 0)].copy()
            if ins_rate or del_rate:
                ev = rng.random(len(seq))
                out = []
                for i, b in enumerate(seq):
                    if ev[i] < del_rate:
                        continue
                    out.append(b)
                    if ev[i] > 1.0 - ins_rate:
                        out.append(acgt[rng.integers(0, 4)])
                seq = np.array(out, np.uint8)
            if rng.random() < indel_reads:
                at, k = int(rng.integers(10, read_len - 10)), int(rng.integers(1, 4))
                if rng.random() < 0.5:
                    seq = np.concatenate([seq[:at], seq[at + k:]])
                else:
                    seq = np.concatenate([seq[:at], acgt[rng.integers(0, 4, k)], seq[at:]])
            seq = seq[:read_len]
            sub = rng.random(read_len) < sub_rate
            seq[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
            s = seq.tobytes()
            strand = "+"
            if rng.random() < 0.5:
                s = s.translate(COMP)[::-1]
                strand = "-"
            f.write(f"@r{r}_{pos}_{strand}\n".encode() + s + b"\n+\n" + b"I" * read_len + b"\n")


def run(which: str, d: Path, threads: int = 4, extra: Sequence[str] = (), out_name: str = "out.sam") -> List[str]:
    """Run one NGM binary; return the sorted SAM body without @PG (command line differs)."""
    env = dict(os.environ)
    ocl = HERE / "_ref" / "ocl"
    env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
    env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    out = d / out_name
    cmd = [str(binary(which)), "-r", str(d / "ref.fa"), "-q", str(d / "reads.fq"), "-o", str(out), "-t", str(threads), "--no-progress", *extra]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=d)
    if p.returncode != 0 or not out.exists():
        raise RuntimeError(f"{which} failed ({p.returncode}):\n{p.stdout[-1500:]}\n{p.stderr[-1500:]}")
    global last_log
    last_log = p.stdout + p.stderr
    lines = [ln for ln in out.read_text().splitlines() if not ln.startswith("@PG")]
    return sorted(lines)


last_log = ""


def logged_sensitivity() -> float:
    """'Estimated sensitivity: %f' of the last run (ReadProvider.cpp:357)."""
    import re
    m = re.search(r"Estimated sensitivity: ([0-9.]+)", last_log)
    if m is None:
        raise RuntimeError("no sensitivity estimate in the log:\n" + last_log[-1500:])
    return float(m.group(1))


def write_paired_inputs(d: Path, ref_len: int = 600_000, n_frags: int = 2_000, read_len: int = 100, seed: int = 20261020, sub_rate: float = 0.02,
                        indel_reads: float = 0.1, insert_mean: float = 400.0, insert_sd: float = 40.0) -> None:
    """Seeded paired-end input (interleaved FASTQ, names <frag>/1 and <frag>/2): two contigs, three copies of a 4 kb segment and a tandem
    duplication (equal pair scores -> the insert-size tie-break of ScoreBuffer::CheckPairs), fragments ~ N(mean, sd) in FR orientation on
    either strand; some fragments have an unmappable mate, mates on different contigs / far apart / on the same strand (PairedFail)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    c1 = acgt[rng.integers(0, 4, ref_len)].copy()
    p1, p2, p3, pt = ref_len // 12, ref_len // 3, (7 * ref_len) // 12, ref_len // 5
    seg = c1[p1: p1 + 4_000].copy()
    c1[p2: p2 + 4_000] = seg
    c1[p2 + 700: p2 + 710] = acgt[rng.integers(0, 4, 10)]           # the copies are not all identical
    c1[p3: p3 + 4_000] = seg
    c1[pt: pt + 600] = c1[pt - 600: pt]                             # tandem duplication, period 600
    c2 = acgt[rng.integers(0, 4, ref_len // 3 + 1)].copy()
    c2[10_000:14_000] = seg
    contigs = [c1, c2]
    with open(d / "ref.fa", "wb") as f:
        for i, c in enumerate(contigs):
            f.write(b">chr%d synthetic\n" % (i + 1))
            for j in range(0, len(c), 60):
                f.write(c[j: j + 60].tobytes() + b"\n")
    hot = [(0, p1), (0, p2), (0, p3), (1, 10_000), (0, pt - 600)]

    def mutate(seq):
        seq = seq.copy()
        if rng.random() < indel_reads:
            at, k = int(rng.integers(10, read_len - 10)), int(rng.integers(1, 4))
            seq = np.concatenate([seq[:at], seq[at + k:]]) if rng.random() < 0.5 else np.concatenate([seq[:at], acgt[rng.integers(0, 4, k)], seq[at:]])
        seq = seq[:read_len]
        sub = rng.random(read_len) < sub_rate
        seq[sub] = acgt[rng.integers(0, 4, int(sub.sum()))]
        return seq.tobytes()

    with open(d / "reads.fq", "wb") as f:
        for i in range(n_frags):
            kind = i % 23
            ci = int(rng.integers(0, 2))
            c = contigs[ci]
            ins = max(read_len + 5, int(rng.normal(insert_mean, insert_sd)))
            if kind in (3, 4, 5, 6, 7):                             # inside / across a repeated segment
                hc, hs = hot[int(rng.integers(0, len(hot)))]
                ci, c = hc, contigs[hc]
                pos = hs + int(rng.integers(-300, 3_900))
            else:
                pos = int(rng.integers(0, len(c) - ins - 8))
            left = mutate(c[pos: pos + read_len + 4])
            right = mutate(c[pos + ins - read_len: pos + ins + 4])
            m1, m2 = left, right[:read_len].translate(COMP)[::-1]    # FR: mate 2 is the reverse complement of the fragment's right end
            if kind == 9:                                           # mates on different contigs
                oc = contigs[1 - ci]
                op = int(rng.integers(0, len(oc) - read_len - 8))
                m2 = mutate(oc[op: op + read_len + 4]).translate(COMP)[::-1]
            if kind == 10:                                          # too far apart
                op = (pos + 5_000) % (len(c) - read_len - 8)
                m2 = mutate(c[op: op + read_len + 4]).translate(COMP)[::-1]
            if kind == 11:                                          # same strand
                m2 = right[:read_len]
            if kind == 12:
                m1 = acgt[rng.integers(0, 4, read_len)].tobytes()   # unmappable first mate
            if kind == 13:
                m2 = acgt[rng.integers(0, 4, read_len)].tobytes()
            if kind == 14:
                m1 = acgt[rng.integers(0, 4, read_len)].tobytes()
                m2 = acgt[rng.integers(0, 4, read_len)].tobytes()
            if rng.random() < 0.5:                                  # the fragment comes from the other strand: the mates swap roles
                m1, m2 = m2, m1
            f.write(f"@f{i}_{ci}_{pos}/1\n".encode() + m1 + b"\n+\n" + b"I" * len(m1) + b"\n")
            f.write(f"@f{i}_{ci}_{pos}/2\n".encode() + m2 + b"\n+\n" + b"I" * len(m2) + b"\n")
