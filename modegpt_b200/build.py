"""Build libmodegpt_b200.so (sm_100a) in-tree with nvcc.

`python -m modegpt_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but travels
to the GPU box with the repo snapshot; nothing is JIT-compiled at run time.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = CSRC / "_obj"
LIB = PKG / "libmodegpt_b200.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src: Path, verbose: bool) -> Path:
    OBJ.mkdir(exist_ok=True)
    obj = OBJ / (src.stem + ".o")
    stamp_file = OBJ / (src.stem + ".stamp")
    stamp = _stamp(src)
    if obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"nvcc failed for {src.name}")
    if verbose:
        sys.stderr.write(r.stderr)
    stamp_file.write_text(stamp)
    return obj


def build_library(verbose: bool = False, force: bool = False) -> Path:
    srcs = _sources()
    if force:
        for f in OBJ.glob("*.stamp"):
            f.unlink()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    p = build_library(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(p)
