"""Build libcoarse3d_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python coarse3d_b200/build.py [--force] [--verbose]     (run as a script: importing
    the package needs the library to exist already)

No torch headers are involved: the library is plain CUDA behind `extern "C"`
(include/coarse3d_b200.h).  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "csrc", "build")
LIB = os.path.join(HERE, "libcoarse3d_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "-Xptxas", "-v"]
# Sources whose float32 arithmetic must match the reference operation by
# operation (no FMA contraction): projection and KNN are bit-exact contracts.
NO_FMA = {"project.cu", "knn.cu"}


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; coarse3d_b200 has no non-CUDA fallback")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/coarse3d_b200.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p) and (f.endswith((".cu", ".cuh", ".h"))):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + COMMON).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every csrc/*.cu and link the shared library.  Returns its path."""
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "digest.txt")
    digest = _digest()
    if (not force and os.path.exists(LIB) and os.path.exists(stamp)
            and open(stamp).read() == digest):
        return LIB
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD, src[:-3] + ".o")
        cmd = [nvcc, *ARCH, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
        if src in NO_FMA:
            cmd.insert(1, "-fmad=false")
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, src[:-3] + ".ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout + r.stderr))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
