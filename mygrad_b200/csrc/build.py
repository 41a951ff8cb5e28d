"""Builds libmgb200.so (sm_100a) in-tree with nvcc.

Run as ``python mygrad_b200/csrc/build.py`` (by path: the package itself refuses to import
without the library) or through ``__graft_entry__.build()``.
Objects are cached by source mtime; ``--force`` rebuilds everything.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE.parent
LIB = PKG / "libmgb200.so"
OBJ_DIR = HERE / "build"

SOURCES = ["runtime.cu", "pool.cu", "elementwise.cu", "conv.cu", "nnops.cu", "batchnorm.cu", "comm.cu", "peaks.cu", "einsum.cu"]

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
        raise RuntimeError("nvcc not found; libmgb200.so cannot be built")
    return exe


def _deps_mtime() -> float:
    hdrs = list(HERE.glob("*.cuh")) + [PKG.parent / "include" / "mgb.h", Path(__file__)]
    return max(h.stat().st_mtime for h in hdrs)


def _compile(src: str, force: bool) -> tuple[str, str]:
    obj = OBJ_DIR / (src + ".o")
    s = HERE / src
    newest = max(s.stat().st_mtime, _deps_mtime())
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return str(obj), ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(s), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return str(obj), r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(exist_ok=True)
    srcs = [s for s in SOURCES if (HERE / s).exists()]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    log = "\n".join(l for _, l in results if l)
    if log:
        (OBJ_DIR / "ptxas.log").write_text(log)
        if verbose:
            print(log)
    newest_obj = max(Path(o).stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest_obj:
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
