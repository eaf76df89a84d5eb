"""Builds mtrl_b200/libmtrl_b200.so from csrc/*.cu with nvcc for sm_100a (in-tree, no JIT cache).

Run as `python -m mtrl_b200.build`; __graft_entry__.build() calls build_library().
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libmtrl_b200.so"
STAMP = PKG / "build" / "stamp.txt"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; mtrl_b200 needs the CUDA toolkit to build its kernels")
    return exe


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "mtrl_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def have_nvcc() -> bool:
    return bool(shutil.which("nvcc")) or os.path.exists("/usr/local/cuda/bin/nvcc")


def is_current() -> bool:
    """True when the shared library was built from exactly the sources, header and flags that are on disk now."""
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _digest()


def build_library(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu to an object (in parallel) and link the shared library.  A no-op when the stamp matches;
    otherwise the build runs under an exclusive file lock (several ranks may import the package at the same time)."""
    if not force and is_current():
        return LIB
    import fcntl

    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    with open(objdir / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():   # another process built it while this one waited
                return LIB
            return _build_locked(objdir, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(objdir: Path, verbose: bool) -> Path:
    digest = _digest()
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name} ====\n{out}")
        if p.returncode != 0:
            failed = True
    (objdir / "nvcc.log").write_text("\n".join(log))
    if verbose or failed:
        print("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see mtrl_b200/build/nvcc.log")
    link = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    subprocess.run(link, check=True)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose=True)
    print("built", path)
