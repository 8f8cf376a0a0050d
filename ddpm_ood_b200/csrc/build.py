"""Build libddpm_ood_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Usage: python -m ddpm_ood_b200.csrc.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
LIB = CSRC / "libddpm_ood_b200.so"
SOURCES = sorted(p.name for p in CSRC.glob("*.cu"))
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def have_nvcc() -> bool:
    import shutil

    return Path(_nvcc()).exists() or shutil.which(_nvcc()) is not None


def _deps():
    return list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((CSRC.parent.parent / "include").glob("*.h"))


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(d.stat().st_mtime > t for d in _deps())


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every .cu and link the shared library. Safe under concurrency (N ranks of one torchrun launch, pytest-xdist
    workers): an exclusive flock serialises builders, whoever gets the lock second finds the library fresh and returns;
    objects and the library are written to temporary names and renamed into place, so a reader never sees a partial
    file."""
    import fcntl
    import tempfile

    if not force and not needs_build():
        return LIB
    lock_path = CSRC / ".build.lock"
    with open(lock_path, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():  # another process built it while we waited
                return LIB
            with tempfile.TemporaryDirectory(prefix=".build_", dir=str(CSRC)) as tmp:
                tmp = Path(tmp)
                objs = []
                procs = []
                for src in SOURCES:
                    obj = CSRC / (src[:-3] + ".o")
                    objs.append(str(obj))
                    # per-file incremental: an object newer than every header and its own source is reused
                    hdr = max(d.stat().st_mtime for d in _deps() if d.suffix != ".cu")
                    if not force and obj.exists() and obj.stat().st_mtime > max(hdr, (CSRC / src).stat().st_mtime):
                        continue
                    part = tmp / obj.name
                    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(part)]
                    if verbose:
                        print(" ".join(cmd), flush=True)
                    procs.append((src, part, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                                    stderr=subprocess.STDOUT, text=True)))
                failed = False
                for src, part, obj, pr in procs:
                    out, _ = pr.communicate()
                    if pr.returncode != 0:
                        failed = True
                        print(f"--- {src} failed ---\n{out}", file=sys.stderr)
                    else:
                        os.replace(part, obj)
                        if verbose and out.strip():
                            print(out)
                if failed:
                    raise RuntimeError("nvcc failed")
                part = tmp / LIB.name
                cmd = [_nvcc(), "-shared", "-o", str(part), *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
                if verbose:
                    print(" ".join(cmd), flush=True)
                subprocess.check_call(cmd)
                os.replace(part, LIB)  # atomic: readers see the old library or the complete new one
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print(LIB)
