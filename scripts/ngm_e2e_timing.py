"""Whole-program timing of the unmodified NextGenMap with its own CPU backend vs the CUDA backend (link seam).
Not the bench metric (that is the hot path): this shows what NGM itself gains while its candidate search (CS) is
still on the host ('host-side CS ... becomes the end-to-end bottleneck')."""
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import ngm_e2e as e2e  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ref_len = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
threads = len(os.sched_getaffinity(0))
with tempfile.TemporaryDirectory(prefix="ngm_t_") as td:
    d = Path(td)
    e2e.write_inputs(d, ref_len=ref_len, n_reads=n_reads, read_len=150)
    env = dict(os.environ)
    ocl = e2e.HERE / "_ref" / "ocl"
    env["OPENCL_VENDOR_PATH"] = str(ocl / "vendor")
    env["LD_LIBRARY_PATH"] = str(ocl / "lib") + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    subprocess.run([str(e2e.binary("ref")), "-r", str(d / "ref.fa"), "-t", str(threads), "--no-progress"], env=env, capture_output=True, cwd=d)  # index only
    res, done = {}, {}
    for which in ("ref", "cuda", "cuda_strict", "ref", "cuda", "cuda_strict"):
        t0 = time.time()
        p = subprocess.run([str(e2e.binary(which)), "-r", str(d / "ref.fa"), "-q", str(d / "reads.fq"), "-o", str(d / f"{which}.sam"), "-t", str(threads),
                            "--no-progress"], env=env, capture_output=True, text=True, cwd=d)
        dt = time.time() - t0
        res.setdefault(which, []).append(dt)
        import re
        m = re.search(r"elapsed: ([0-9.]+)s", p.stdout + p.stderr)
        done.setdefault(which, []).append(float(m.group(1)) if m else None)
    same = sorted(l for l in open(d / "ref.sam") if not l.startswith("@PG")) == sorted(l for l in open(d / "cuda.sam") if not l.startswith("@PG"))
print(json.dumps({"reads": n_reads, "ref_len": ref_len, "threads": threads, "wall_s": res, "ngm_done_elapsed_s": done, "reads_per_s": {k: n_reads / min(v) for k, v in res.items()},
                  "sam_identical": same}))
