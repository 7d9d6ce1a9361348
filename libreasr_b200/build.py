"""Build recipe for the CUDA library (sm_100a only).  `python -m libreasr_b200.build`.

Compiles every .cu under libreasr_b200/csrc with nvcc into ONE shared object,
libreasr_b200/lib/librnnt_b200.so, exporting the C ABI of include/rnnt_b200.h.  The
.so is built in-tree (git-ignored) so that it travels to the GPU box with the snapshot.
"""
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librnnt_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
        os.path.join(ROOT, "include", "rnnt_b200.h"), os.path.abspath(__file__)]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False, extra_flags=()):
    """Compile (if stale) and return the path of the shared library."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lcuda"]
    subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True,
                extra_flags=("-Xptxas", "-v") if "--ptxas" in sys.argv else ()))
