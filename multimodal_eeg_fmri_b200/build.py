"""In-tree build of the CUDA C-ABI library (sm_100a only).

`python -m multimodal_eeg_fmri_b200.build` or `build()` compiles csrc/*.cu with nvcc into
multimodal_eeg_fmri_b200/_C/libxmodal_b200.so.  The .so is git-ignored but travels to the GPU box
with the repo snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OUT_DIR = PKG_DIR / "_C"
LIB_PATH = OUT_DIR / "libxmodal_b200.so"
STAMP = OUT_DIR / "build.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
]
LINK_FLAGS = ["-cudart", "static", "-shared"]
OBJ_DIR = PKG_DIR.parent / "build" / "obj"  # git-ignored; per-file objects keyed by a hash of source + headers + flags


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "xmodal_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    return h.hexdigest()


def _object_key(src: Path, extra) -> str:
    h = hashlib.sha256(src.read_bytes())
    for p in sorted(list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "xmodal_b200.h"]):
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()[:20]


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == _fingerprint()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile the library if sources changed; returns the .so path."""
    if not force and is_current():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    OUT_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    extra = os.environ.get("XM_NVCC_FLAGS", "").split()  # e.g. -DXM_FA_TRACE (with --force)
    if verbose:
        extra += ["-Xptxas", "-v"]

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / f"{src.stem}.{_object_key(src, extra)}.o"
        if obj.exists() and not force and not verbose:
            return obj
        res = subprocess.run([_nvcc(), *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)], capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {src.name}")
        if verbose:
            sys.stderr.write(f"== {src.name}\n" + res.stdout + res.stderr)
        for old in OBJ_DIR.glob(f"{src.stem}.*.o"):  # drop objects of earlier versions of this file
            if old != obj:
                old.unlink()
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:  # one nvcc per translation unit
        objs = list(pool.map(compile_one, _sources()))
    res = subprocess.run([_nvcc(), *NVCC_FLAGS, *LINK_FLAGS, "-o", str(LIB_PATH), *map(str, objs)], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libxmodal_b200.so")
    STAMP.write_text(_fingerprint())
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
