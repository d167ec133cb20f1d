"""Build liblatentknn.so in-tree with nvcc for sm_100a.

    python -m latent_rag_b200.build [--force]

The shared library lands next to this file (git-ignored, but it travels to the GPU box
with the repo snapshot).  Objects are cached under latent_rag_b200/csrc/_obj and rebuilt
when a source or header is newer.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from typing import List

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(PKG, "liblatentknn.so")

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-Xptxas",
    "-v",
    "-I",
    os.path.join(ROOT, "include"),
    "-I",
    CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblatentknn.so cannot be built (there is no CPU fallback)")


def _sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> List[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(ROOT, "include", "latentknn.h"))
    return hs


def _compile(nvcc: str, src: str, obj: str, log_dir: str) -> str:
    extra = os.environ.get("LK_NVCC_EXTRA", "").split()  # experiments, e.g. -DLK_EPI_WARPS=16
    cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(log_dir, os.path.basename(src) + ".ptxas.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in _headers())
    jobs, objs = [], []
    for src in _sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header)
        if stale:
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print(f"[latent_rag_b200.build] nvcc sm_100a: {', '.join(os.path.basename(s) for s, _ in jobs)}", flush=True)
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda so: _compile(nvcc, so[0], so[1], OBJ), jobs))
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"[latent_rag_b200.build] linked {LIB}", flush=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
