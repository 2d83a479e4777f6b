"""Build recipe for the in-tree CUDA library (sm_100a only).

``python -m d2r_b200.build`` (or ``__graft_entry__.build()``) compiles every ``csrc/*.cu``
with ``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``d2r_b200/csrc/libd2r_b200.so``.  The .so stays in-tree (git-ignored) so that it travels to
the GPU box with the snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "libd2r_b200.so")
SOURCES = ["c_api.cu", "gemm_tc.cu", "attn_fused.cu", "gemm_simt.cu", "elementwise.cu", "router.cu", "aggregate.cu", "saf.cu"]
HEADERS = ["common.cuh", "ptx.cuh", "gemm_tc_epilogue.cuh", "gemm_tc2.cuh", "tc_epi_common.cuh", "tc_host.cuh", os.path.join("..", "..", "include", "d2r_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _flags():
    """NVCC_FLAGS plus optional extra flags from the environment (D2R_NVCC_EXTRA, e.g. '-DD2R_MBAR_SPIN_LOG2=22')."""
    return NVCC_FLAGS + os.environ.get("D2R_NVCC_EXTRA", "").split()


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(_flags()).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
    stamp = obj + ".sha"
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    dig = _digest(deps)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [nvcc] + _flags() + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="--force" in sys.argv))
