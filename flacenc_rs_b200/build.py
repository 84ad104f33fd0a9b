"""Builds libflacenc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

fb_api.cu holds the C ABI, the host pipeline and the kernels that do not depend on the LPC tap-window size;
fb_inst.cu holds the ones that do and is compiled once per window size G (-DFB_INST_G=...), all objects in
parallel, then linked into one shared library."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libflacenc_b200.so")
RINGS = [4, 8, 12, 16, 20, 24]
HEADERS = ["fb_common.h", "fb_kernels.cuh", "fb_fused.cuh", "fb_host.h", "fb_launch.h", "fb_md5.h",
           os.path.join("..", "..", "include", "flacenc_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no implicit contraction: every fused multiply-add is an explicit __fma_rn/__fmaf_rn exactly
    # where the reference calls mul_add, so device floats are bit-identical to the scalar CPU order
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC",
]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _units():
    """(source, object, extra flags) of every translation unit"""
    units = [("fb_api.cu", os.path.join(OBJ, "fb_api.o"), [])]
    for g in RINGS:
        units.append(("fb_inst.cu", os.path.join(OBJ, f"fb_inst_g{g}.o"), [f"-DFB_INST_G={g}"]))
    return units


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in ["fb_api.cu", "fb_inst.cu"] + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = nvcc_path()

    def compile_one(unit):
        src, obj, extra = unit
        extra = extra + os.environ.get("FB200_NVCC_EXTRA", "").split()  # experiments: e.g. -DFB_KAP_MAXNREG=96
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        return subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(len(RINGS) + 1, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, _units()))
    for res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [u[1] for u in _units()]
    res = subprocess.run(link, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
