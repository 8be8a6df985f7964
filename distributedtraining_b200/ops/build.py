"""In-tree build of the sm_100a kernel library.

Every ``csrc/*.cu`` file is compiled with plain ``nvcc`` (no torch headers -> seconds per file) into ONE shared
object ``distributedtraining_b200/build/libdtb200.so`` exposing a C ABI that :mod:`distributedtraining_b200.ops._lib`
binds with ``ctypes``.  nvcc cross-compiles without a GPU, so this runs on the CPU-only dev box; the resulting ``.so``
travels to the GPU box with the repo snapshot.

The reference has no native code at all -- this library is the new surface that replaces every
implicit PyTorch/cuBLAS kernel on its hot paths.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent.parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "build"
LIB_PATH = BUILD_DIR / "libdtb200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(path: Path) -> str:
    h = hashlib.sha256()
    h.update(path.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = BUILD_DIR / (src.stem + ".o")
    stamp = BUILD_DIR / (src.stem + ".stamp")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp.write_text(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    """Compile all kernels for sm_100a and link ``libdtb200.so`` in-tree.  Incremental (content-hash stamps)."""
    BUILD_DIR.mkdir(exist_ok=True)
    if force:
        for f in BUILD_DIR.glob("*.stamp"):
            f.unlink()
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs) or 1)) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    link_stamp = BUILD_DIR / "link.stamp"
    want = "|".join((BUILD_DIR / (s.stem + ".stamp")).read_text() for s in srcs)
    if LIB_PATH.exists() and link_stamp.exists() and link_stamp.read_text() == want:
        return LIB_PATH
    cmd = [_nvcc(), "-shared", "-o", str(LIB_PATH), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    link_stamp.write_text(want)
    return LIB_PATH


def is_built() -> bool:
    return LIB_PATH.exists()


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
