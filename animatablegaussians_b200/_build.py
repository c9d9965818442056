"""In-tree build of the sm_100a CUDA library (nvcc cross-compiles without a GPU).

The product is ONE shared library with a plain C ABI (include/*.h):
    animatablegaussians_b200/libagr_b200.so
built from animatablegaussians_b200/csrc/*.cu.  It is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
import glob
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libagr_b200.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG_DIR, "..", "include", "*.h"))


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def build_extension(force=False, verbose=True):
    """Compile every .cu under csrc/ for sm_100a into libagr_b200.so (parallel per file)."""
    if not force and not is_stale():
        return LIB_PATH
    objdir = os.path.join(PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    newest_hdr = max([os.path.getmtime(f) for f in _deps() if not f.endswith(".cu")] or [0])
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), newest_hdr):
            continue
        cmd = ["nvcc", *NVCC_FLAGS, "-I" + os.path.join(PKG_DIR, "..", "include"), "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError("nvcc failed for %s" % src)
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH, *objs, "-lcuda"]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    build_extension(force="--force" in sys.argv)
