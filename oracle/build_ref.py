"""TEST INFRASTRUCTURE — builds the UNMODIFIED reference kernels into oracle/_ref/.

The reference sources are compiled *where they lie* under /root/reference (nothing is
copied into this repo); only build outputs land in oracle/_ref/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).  Only tests/, __graft_entry__.smoke()
and bench.py's reference / cpu_baseline legs may load what this script builds.

Built artefacts
  oracle/_ref/libref_rasterizer.so   reference CUDA rasterizer (forward.cu, backward.cu,
                                     rasterizer_impl.cu) + oracle/ref_shim.cu (C-ABI)
  oracle/_ref/ref_fused.so           reference `fused` torch extension (fused_bias_act*)
  oracle/_ref/ref_upfirdn2d.so       reference `upfirdn2d` torch extension

`-include cstdint` works around the missing <cstdint> include in the reference's
cuda_rasterizer/rasterizer_impl.h with GCC 13.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AGR_REFERENCE_ROOT", "/root/reference")
RAST = os.path.join(REF_ROOT, "gaussians", "diff_gaussian_rasterization_depth_alpha")
SUNET = os.path.join(REF_ROOT, "network", "styleunet")
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_rasterizer(force=False):
    srcs = [os.path.join(RAST, "cuda_rasterizer", f) for f in ("forward.cu", "backward.cu", "rasterizer_impl.cu")]
    srcs.append(os.path.join(HERE, "ref_shim.cu"))
    out = os.path.join(OUT, "libref_rasterizer.so")
    if not force and _newer(out, srcs):
        return out
    os.makedirs(OUT, exist_ok=True)
    _run(["nvcc", "-O3", "-std=c++17", *ARCH, "-include", "cstdint", "-Xcompiler", "-fPIC", "-shared",
          "-I" + os.path.join(RAST, "third_party", "glm"), "-I" + RAST, "-o", out, *srcs])
    return out


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = ["-I" + p for p in ce.include_paths()] + ["-I" + sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    libs = ["-L" + libdir, "-Wl,-rpath," + libdir, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
            "-lc10_cuda", "-ltorch_cuda", "-L/usr/local/cuda/lib64", "-lcudart"]
    abi = "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, libs, abi


def build_styleunet_op(name, cpp, cu, force=False):
    """Reference network/styleunet/setup.py builds `fused` and `upfirdn2d` as two torch
    extension modules; same sources, module names prefixed ref_ so they cannot shadow
    anything in the product."""
    srcs = [os.path.join(SUNET, cpp), os.path.join(SUNET, cu)]
    mod = "ref_" + name
    out = os.path.join(OUT, mod + ".so")
    if not force and _newer(out, srcs):
        return out
    os.makedirs(OUT, exist_ok=True)
    inc, libs, abi = _torch_flags()
    common = ["-DTORCH_EXTENSION_NAME=" + mod, "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    obj_cpp = os.path.join(OUT, mod + "_host.o")
    obj_cu = os.path.join(OUT, mod + "_dev.o")
    _run(["g++", "-O2", "-std=c++17", "-fPIC", *common, *inc, "-c", srcs[0], "-o", obj_cpp])
    _run(["nvcc", "-O3", "-std=c++17", *ARCH, "-Xcompiler", "-fPIC", *common, *inc,
          "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
          "--expt-relaxed-constexpr", "-c", srcs[1], "-o", obj_cu])
    _run(["g++", "-shared", obj_cpp, obj_cu, "-o", out, *libs])
    return out


def build_all(force=False):
    if not os.path.isdir(REF_ROOT):
        print("reference tree not present (%s): keeping prebuilt oracle/_ref" % REF_ROOT)
        return False
    build_rasterizer(force)
    try:
        build_styleunet_op("fused", "fused_bias_act.cpp", "fused_bias_act_kernel.cu", force)
        build_styleunet_op("upfirdn2d", "upfirdn2d.cpp", "upfirdn2d_kernel.cu", force)
    except Exception as e:  # the styleunet ops are only needed by the reference arm
        print("WARNING: reference styleunet ops failed to build:", e)
    return True


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
