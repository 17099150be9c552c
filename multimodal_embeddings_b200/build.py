"""Builds libpagegeom.so (the C-ABI library declared in include/pagegeom.h) in-tree with nvcc.

    python -m multimodal_embeddings_b200.build [--force] [--verbose]

sm_100a only; -fmad=false so the fp64 box arithmetic is never contracted into FMAs;
-lineinfo so ncu's source page maps back to the .cu files.

Staleness is decided by a content hash of the sources (stored next to the library), not by
mtimes — a snapshot copied to another machine keeps the prebuilt library.  Concurrent callers
(one process per GPU under torchrun) are serialised by a file lock and the library is replaced
atomically.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpagegeom.so")
STAMP = LIB + ".srchash"
SOURCES = ["pg_tiler.cu", "pg_boxes.cu", "pg_json.cu", "pg_comm.cu", "pg_jpeg.cu"]
HEADERS = ["pg_common.cuh", "pg_math.h", "pg_fmt.h", "pg_ryu_tables.h", "pg_jpeg.h", os.path.join("..", "..", "include", "pagegeom.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "--threads", "0", "-ldl",
]


def source_hash() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def _stale(want: str) -> bool:
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != want


def build(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if not force and not _stale(want):
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale(want):  # another rank built it while we waited
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-o", tmp] + [os.path.join(CSRC, f) for f in SOURCES]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed building libpagegeom.so")
            if verbose:
                sys.stderr.write(res.stderr)
            os.replace(tmp, LIB)
            with open(STAMP, "w") as f:
                f.write(want + "\n")
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def build_variant(name: str, extra_flags) -> str:
    """A second build of the same sources with extra nvcc flags, written to _variants/libpagegeom_<name>.so and loaded
    through PAGEGEOM_LIB (A/B timing builds; `checked`: -DPG_CHECKED, device-side bounds assertions — pg_common.cuh)."""
    out_dir = os.path.join(HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libpagegeom_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-o", out] + [os.path.join(CSRC, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building {out}")
    return out


if __name__ == "__main__":
    if "--checked" in sys.argv:
        print(build_variant("checked", ["-DPG_CHECKED"]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
