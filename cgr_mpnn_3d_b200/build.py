"""In-tree build of libcgr_b200.so (nvcc, sm_100a only).

``python -m cgr_mpnn_3d_b200.build`` or ``__graft_entry__.build()``.  The shared library is
written next to this file so it travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcgr_b200.so")
STAMP = os.path.join(HERE, ".libcgr_b200.stamp")
SOURCES = ["api.cu", "collate.cu", "simt.cu", "tc.cu", "optim.cu", "featurize.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
if os.environ.get("CGR_FWD_STAMPS"):      # debug build: clock64 phase stamps in the fused forward kernel (tools/fwd_phase_timing.py)
    NVCC_FLAGS.append("-DCGR_FWD_STAMPS")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libcgr_b200.so")


def _digest() -> str:
    h = hashlib.sha256()
    root = os.path.dirname(HERE)
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(root, "include", "cgr_b200.h")]
    for f in files:
        with open(f, "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``libcgr_b200.so``; returns its path."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(src):
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        return src, obj, subprocess.run(cmd, capture_output=True, text=True)

    objs = []
    logs = []
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as pool:     # one nvcc per source file
        for src, obj, r in pool.map(compile_one, SOURCES):
            logs.append(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            objs.append(obj)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link of libcgr_b200.so failed")
    with open(os.path.join(HERE, "ptxas_info.log"), "w") as fh:
        fh.write("\n".join(logs))
    with open(STAMP, "w") as fh:
        fh.write(digest)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
