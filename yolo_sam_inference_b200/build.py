"""Builds libysi.so (hand-written sm_100a CUDA + the C ABI of include/ysi.h) in-tree with nvcc.

The library is plain CUDA runtime code (no torch types anywhere); nvcc cross-compiles it without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libysi.so")
# One library per 16-bit operand encoding of the tensor-core contractions (csrc/common.h): same sources, same kernels,
# only the conversion instructions and the tcgen05 instruction-descriptor format bits differ.
VARIANTS = {"bf16": ("libysi.so", []), "fp16": ("libysi_fp16.so", ["-DYSI_OP_FP16=1"])}
SOURCES = ["gemm.cu", "attn.cu", "encoder.cu", "decoder.cu", "postproc.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
              "-diag-suppress", "177"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def lib_path(precision: str = "bf16") -> str:
    return os.path.join(HERE, VARIANTS[precision][0])


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ysi.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, precision: str = "all") -> str:
    """Compile every .cu for sm_100a and link the libraries next to this file. Returns the bf16 library path
    (or the requested variant's)."""
    if precision == "all":
        for v in VARIANTS:
            build(force, verbose, v)
        return LIB
    lib, defines = lib_path(precision), VARIANTS[precision][1] + os.environ.get("YSI_NVCC_DEFINES", "").split()
    if not force and not _stale(lib):
        return lib
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build", precision)
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xlinker", "-Bsymbolic"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    prec = "all"
    for a in sys.argv[1:]:
        if a.startswith("--precision="):
            prec = a.split("=", 1)[1]
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, precision=prec))
