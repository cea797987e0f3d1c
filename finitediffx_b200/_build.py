"""In-tree nvcc build of libfdx_b200.so (sm_100a only).

``python -m finitediffx_b200._build`` or ``finitediffx_b200._build.build()``.  The shared object
lands in ``finitediffx_b200/_lib/`` so it travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIBDIR, "libfdx_b200.so")
SOURCES = ("fdx_plan.cu", "fdx_axis.cu", "fdx_rowtma.cu", "fdx_coltma.cu", "fdx_dev_alpha.cu", "fdx_fused2d.cu", "fdx_hessian3d.cu", "fdx_fused3d_lap.cu", "fdx_fused3d_div.cu", "fdx_fused3d_grad.cu", "fdx_fused3d_curl.cu")
NVCC_FLAGS = [
    *(["-DFDX_SWEEP"] if os.environ.get("FDX_SWEEP") else []),
    *([f"-DFDX_SWEEP_P={os.environ['FDX_SWEEP_P']}"] if os.environ.get("FDX_SWEEP_P") else []),
    *([f"-DFDX_SWEEP_TSIZE={os.environ['FDX_SWEEP_TSIZE']}"] if os.environ.get("FDX_SWEEP_TSIZE") else []),
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the fdx_b200 CUDA library cannot be built")


def _digest() -> str:
    h = hashlib.sha1()
    files = sorted(os.listdir(CSRC)) + ["../../include/fdx_b200.h"]
    for name in files:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            with open(path, "rb") as f:
                h.update(name.encode() + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha1")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
