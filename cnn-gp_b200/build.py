"""Build libcnngp.so (sm_100a) in-tree with nvcc.  Usage: python cnn-gp_b200/build.py [--force]

The shared library is the whole native product: hand-written CUDA kernels plus the C ABI of
include/cnngp.h.  It is built next to this file so that it travels with the repository
snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcnngp.so")
SOURCES = ["plan.cu", "gram_generic.cu", "gram_variance.cu", "gram_fused.cu", "gram_fnet.cu", "chol_f64.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-ccbin", "/usr/bin/g++"]


def _newer_than(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "cnngp.h"))
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _newer_than(o, [s] + headers):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for msg in ex.map(run, jobs):
            if verbose and msg:
                print(msg)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _newer_than(LIB, objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lcudart"])
    # the HDF5 block store: host code only, its own library (include/cnngp_h5.h)
    h5_src, h5_lib = os.path.join(CSRC, "h5store.cpp"), os.path.join(HERE, "libcnngp_h5.so")
    if force or _newer_than(h5_lib, [h5_src, os.path.join(HERE, "..", "include", "cnngp_h5.h")]):
        run(["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-fPIC", "-shared", "-pthread", h5_src, "-o", h5_lib])
    # measurement-only probes (roofline denominators for bench.py)
    mb_src, mb_lib = os.path.join(CSRC, "microbench.cu"), os.path.join(HERE, "libcnngp_bench.so")
    if force or _newer_than(mb_lib, [mb_src]):
        run([NVCC] + FLAGS + ["-shared", mb_src, "-o", mb_lib, "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
