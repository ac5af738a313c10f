"""Build libsfgwas_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsfgwas_b200.so")
OBJ = os.path.join(HERE, "build")

CU = ["ctx.cu", "hostio.cu", "kernels_ntt.cu", "kernels_encode.cu", "kernels_mactc.cu", "kernels_ks.cu", "kernels_ctalg.cu", "kernels_geno.cu", "kernels_cachefile.cu", "kernels_refresh.cu", "matmult.cu", "capi.cu"]
CPP = ["hostmath.cpp", "cmfile.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _newer(src_paths, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_paths)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(HERE, "..", "include", "sfgwas_b200.h"))
    return hs


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for f in CU:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        if force or _newer([src] + hdrs, obj):
            jobs.append((["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj], obj))
    for f in CPP:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f + ".o")
        if force or _newer([src] + hdrs, obj):
            jobs.append((["g++", "-O2", "-std=gnu++17", "-fPIC", "-c", src, "-o", obj], obj))

    def run(job):
        cmd, obj = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout + r.stderr))
        if verbose:
            print(" ".join(cmd))

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(OBJ, f + ".o") for f in CU + CPP]
    if force or jobs or not os.path.exists(LIB):
        cmd = ["nvcc", "-shared", "-o", LIB] + objs + ["-lquadmath"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout + r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
