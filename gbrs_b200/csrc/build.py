"""Build libgbrs_em.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and `python -m gbrs_b200.csrc.build`.

The shared library is self-contained (CUDA runtime linked statically) and exposes only the C ABI declared in
include/gbrs_em.h; Python loads it with ctypes (gbrs_b200/_lib.py)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_C")
LIB = os.path.join(OUT_DIR, "libgbrs_em.so")
SOURCES = [os.path.join(HERE, "em_kernels.cu"), os.path.join(HERE, "ec_kernels.cu"), os.path.join(HERE, "hmm_kernels.cu"),
           os.path.join(HERE, "gpu_pack.cu"),
           os.path.join(HERE, "pack.cpp"), os.path.join(HERE, "tile_pack.cpp"),
           os.path.join(HERE, "report.cpp")]
HEADERS = [os.path.join(ROOT, "include", "gbrs_em.h"), os.path.join(HERE, "pack_internal.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA path cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS + [os.path.abspath(__file__)])


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    """`out`: alternative output path (tuning builds, loaded through GBRS_LIB_PATH)."""
    if out is not None:
        global LIB
        saved, LIB = LIB, out
        try:
            return build(force=True, verbose=verbose)
        finally:
            LIB = saved
    if not force and not is_stale():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC,-fopenmp,-O3", "-Xptxas", "-v" if verbose else "-O3",
           "-I", os.path.join(ROOT, "include"), "-cudart", "static", "-o", LIB + ".tmp"] + SOURCES + ["-lgomp", "-ldl"]
    for knob in ("GBRS_THREADS", "GBRS_COL_MINBLOCKS", "GBRS_ROW_MINBLOCKS", "GBRS_TILE_THREADS", "GBRS_TILE_MINBLOCKS"):  # tuning experiments only
        if os.environ.get(knob):
            cmd.insert(1, f"-D{knob}=" + os.environ[knob])
    for flag in os.environ.get("GBRS_DEFINES", "").split():  # tuning experiments only: bare -D flags
        cmd.insert(1, "-D" + flag)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
