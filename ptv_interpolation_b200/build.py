"""Build libptvb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m ptv_interpolation_b200.build [--force] [--verbose]

Every translation unit is compiled to its own object file (in parallel, rebuilt only when the
source or a shared header changed) and the objects are linked into the shared library.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libptvb200.so")
STAMP = LIB + ".stamp"
SOURCES = ["cabi.cu", "hash_build.cu", "knn_interp.cu", "knn_stream.cu", "knn_duo.cu", "knn_dispatch.cu",
           "delaunay_linear.cu", "grid_ops.cu", "stencil_fused.cu", "strain_bulk.cu", "projection.cu"]
HEADERS = [os.path.join(CSRC, "ptv_internal.cuh"), os.path.join(CSRC, "knn_common.cuh"), os.path.join(CSRC, "bulk_pipe.cuh"), os.path.join(ROOT, "include", "ptv_b200.h")]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def nvcc_path() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libptvb200.so cannot be built (there is no CPU fallback)")
    return cand


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _obj_stale(src: str) -> bool:
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, src)] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def source_hash() -> str:
    """Content hash of everything the library is built from (sources, headers, flags).  The build
    writes it next to the .so; a stale binary (edited sources, older ABI) is detected by content, so
    the check also works on a copy of the tree whose mtimes were not preserved."""
    h = hashlib.sha256(" ".join(FLAGS[:8]).encode())
    for d in [os.path.join(CSRC, s) for s in _sources()] + HEADERS:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def is_stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:  # ranks of one job must not build concurrently
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not is_stale():
            return LIB
        return _build_locked(force, verbose)


def _build_locked(force: bool, verbose: bool) -> str:
    nvcc = nvcc_path()
    todo = [s for s in _sources() if force or _obj_stale(s)]

    def compile_one(src):
        cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 1) or 1) as ex:
        results = list(ex.map(compile_one, todo))
    failed = False
    for src, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(f"---- {src}\n" + res.stdout + res.stderr)
        failed |= res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libptvb200.so")
    tmp = LIB + ".tmp"
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] +
                         [_obj(s) for s in _sources()] + ["-o", tmp], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libptvb200.so")
    os.replace(tmp, LIB)
    with open(STAMP, "w") as f:
        f.write(source_hash() + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
