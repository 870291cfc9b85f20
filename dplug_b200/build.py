"""Builds libb2pbr.so (the CUDA library behind include/b2pbr.h) in-tree for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only container; the resulting
dplug_b200/libb2pbr.so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb2pbr.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-ftz=false", "-prec-div=true", "-prec-sqrt=true"]

# (source, extra flags).  The integer mip kernels and the exact lighting kernel must not contract a*b+c
# into FMA (the reference's SSE code has none); pbr_tex.cu is free to.
UNITS = [
    ("b2pbr.cu", ["-fmad=false"]),
    ("mip_kernels.cu", ["-fmad=false"] + ["-D%s=%s" % (k, os.environ[k]) for k in ("B2_TMA_ROWS", "B2_TMA_STAGES") if k in os.environ]),
    ("mip_wave.cu", ["-fmad=false"] + ["-D%s=%s" % (k, os.environ[k]) for k in ("B2_WAVE_MIN_CTAS", "B2_WAVE_UNROLL", "B2_WAVE_ITEM_TEXELS") if k in os.environ]),
    ("mip_fused.cu", ["-fmad=false"] + (["-DB2_MIP_DEBUG"] if os.environ.get("B2_MIP_DEBUG") else [])),
    ("pbr_exact.cu", ["-fmad=false"]),
    ("resizer.cu", ["-fmad=false"]),
    ("pbr_tex.cu", ["-DB2_TEX_MIN_CTAS=%s" % os.environ.get("B2_TEX_MIN_CTAS", "5")] + ["-D%s=%s" % (k, os.environ[k]) for k in ("B2_PAIR_MIN_CTAS",) if k in os.environ]),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".inc", ".inl"))]
    headers.append(os.path.join(HERE, "..", "include", "b2pbr.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        subprocess.run([nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"], check=True)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
