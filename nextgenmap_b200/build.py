"""Build the CUDA shared library in-tree: nextgenmap_b200/libngm_b200.so (sm_100a only).

    python -m nextgenmap_b200.build [--force]

nvcc cross-compiles without a GPU; translation units are built in parallel.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
# A/B builds: NGM_B200_DEFINES="-DNGM_FWD_ROW_UNROLL=8" NGM_B200_VARIANT=u8 python -m nextgenmap_b200.build  ->  libngm_b200_u8.so
# (objects under build_u8/); load it with NGM_B200_LIB=<path>.  The product is the default build.
VARIANT = os.environ.get("NGM_B200_VARIANT", "")
EXTRA_DEFINES = os.environ.get("NGM_B200_DEFINES", "").split()
OBJ = PKG / ("build_" + VARIANT if VARIANT else "build")
LIB = PKG / ("libngm_b200_" + VARIANT + ".so" if VARIANT else "libngm_b200.so")
NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def compile_one(src: Path, force: bool) -> Path:
    obj = OBJ / (src.name + ".o")
    deps = [src] + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG.parent / "include").glob("*.h"))
    if force or stale(obj, deps):
        cmd = [NVCC] + ARCH + CUFLAGS + EXTRA_DEFINES + ["-DNGM_HAVE_S16" if (CSRC / "k_score_s16.cu").exists() else "-DNGM_NO_S16",
                                         "-I", str(PKG.parent / "include"), "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = sources()
    deps = srcs + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list((PKG.parent / "include").glob("*.h"))
    if not force and not stale(LIB, deps):          # up to date (the object files do not travel to the GPU box)
        if verbose:
            print(f"up to date: {LIB}")
        return LIB
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: compile_one(s, force), srcs))
    if force or stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-Xlinker", "--exclude-libs,ALL"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
