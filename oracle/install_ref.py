"""TEST / BENCH INFRASTRUCTURE — puts the UNMODIFIED reference where the GPU box can run it through its own code path.

    python -m oracle.install_ref        (run in the build container, where /root/reference exists)

baseline/_ref/                              git-ignored, NOT gpurun-ignored: travels to the GPU box with the snapshot
    AnimatableGaussians/                    verbatim copies of the reference's Python files (config.py, network/,
                                            gaussians/*.py, utils/*.py, smplx/) — nothing is edited
    ext/diff_gaussian_rasterization_depth_alpha/{__init__.py (verbatim), _C.so}
    ext/fused.so, ext/upfirdn2d.so          the reference's three native torch extensions, compiled here from the
                                            sources where they lie under /root/reference (ext.cpp, rasterize_points.cu,
                                            cuda_rasterizer/*.cu; fused_bias_act*.{cpp,cu}; upfirdn2d*.{cpp,cu}) for sm_100a,
                                            under their OWN module names, so `gaussians/gaussian_renderer.py:14`,
                                            `network/styleunet/fused_act.py:30` and `upfirdn2d.py:30` import them unchanged.
The reference has no setup.py at its root (it is a research code base run from its checkout), so the contract's
`pip install --target baseline/_ref /root/reference` has nothing to install; this script is the equivalent recipe.
Third-party packages the reference imports but this image lacks (pytorch3d, plyfile) are stood in by oracle/shims/.
Only tests/, bench.py --impl reference and oracle/ref_stock.py use what this script produces; the product never does."""
import os
import shutil
import sys

from oracle import build_ref as B

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DEST = os.path.join(ROOT, "baseline", "_ref")
PY_DEST = os.path.join(DEST, "AnimatableGaussians")
EXT_DEST = os.path.join(DEST, "ext")

PY_TREES = ["config.py", "network", "gaussians", "utils", "smplx", "base_trainer.py", "main_avatar.py", "dataset", "configs"]
SKIP_DIRS = {"diff_gaussian_rasterization_depth_alpha", "posevocab_custom_ops", "root_finding", "renderer", "weights", "__pycache__"}
KEEP_EXT = {".py", ".yaml", ".yml"}


def copy_python():
    for item in PY_TREES:
        src = os.path.join(B.REF_ROOT, item)
        if os.path.isfile(src):
            os.makedirs(PY_DEST, exist_ok=True)
            shutil.copy2(src, os.path.join(PY_DEST, item))
            continue
        for d, dirs, files in os.walk(src):
            dirs[:] = [x for x in dirs if x not in SKIP_DIRS]
            rel = os.path.relpath(d, B.REF_ROOT)
            for f in files:
                if os.path.splitext(f)[1] in KEEP_EXT:
                    os.makedirs(os.path.join(PY_DEST, rel), exist_ok=True)
                    shutil.copy2(os.path.join(d, f), os.path.join(PY_DEST, rel, f))
    # the LPIPS "lin" weights the reference ships next to its code (7 KB of data; `weights` directories are otherwise skipped)
    lin = os.path.join(B.REF_ROOT, "network", "lpips", "weights", "v0.1", "vgg.pth")
    if os.path.exists(lin):
        os.makedirs(os.path.join(PY_DEST, "network", "lpips", "weights", "v0.1"), exist_ok=True)
        shutil.copy2(lin, os.path.join(PY_DEST, "network", "lpips", "weights", "v0.1", "vgg.pth"))
    pkg = os.path.join(EXT_DEST, "diff_gaussian_rasterization_depth_alpha")
    os.makedirs(pkg, exist_ok=True)
    shutil.copy2(os.path.join(B.RAST, "diff_gaussian_rasterization_depth_alpha", "__init__.py"), os.path.join(pkg, "__init__.py"))


def build_ext(mod, sources, out, includes=(), force=False):
    """One torch extension module `mod` from reference sources (host .cpp with g++, .cu with nvcc)."""
    if not force and B._newer(out, sources):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    inc, libs, abi = B._torch_flags()
    inc = list(inc) + ["-I" + i for i in includes]
    common = ["-DTORCH_EXTENSION_NAME=" + mod, "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    objs = []
    for i, src in enumerate(sources):
        obj = out + ".%d.o" % i
        if src.endswith(".cu"):
            B._run(["nvcc", "-O3", "-std=c++17", *B.ARCH, "-include", "cstdint", "-Xcompiler", "-fPIC", *common, *inc,
                    "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
                    "--expt-relaxed-constexpr", "-c", src, "-o", obj])
        else:
            B._run(["g++", "-O2", "-std=c++17", "-fPIC", "-include", "cstdint", *common, *inc, "-c", src, "-o", obj])
        objs.append(obj)
    B._run(["g++", "-shared", *objs, "-o", out, *libs])
    for o in objs:
        os.remove(o)
    return out


def build_native(force=False):
    cr = os.path.join(B.RAST, "cuda_rasterizer")
    build_ext("_C", [os.path.join(B.RAST, "ext.cpp"), os.path.join(B.RAST, "rasterize_points.cu"), os.path.join(cr, "forward.cu"),
                     os.path.join(cr, "backward.cu"), os.path.join(cr, "rasterizer_impl.cu")],
              os.path.join(EXT_DEST, "diff_gaussian_rasterization_depth_alpha", "_C.so"),
              includes=[os.path.join(B.RAST, "third_party", "glm"), B.RAST], force=force)
    build_ext("fused", [os.path.join(B.SUNET, "fused_bias_act.cpp"), os.path.join(B.SUNET, "fused_bias_act_kernel.cu")],
              os.path.join(EXT_DEST, "fused.so"), force=force)
    build_ext("upfirdn2d", [os.path.join(B.SUNET, "upfirdn2d.cpp"), os.path.join(B.SUNET, "upfirdn2d_kernel.cu")],
              os.path.join(EXT_DEST, "upfirdn2d.so"), force=force)


def install(force=False):
    if not os.path.isdir(B.REF_ROOT):
        print("reference tree not present (%s): keeping the installed baseline/_ref" % B.REF_ROOT)
        return os.path.isdir(PY_DEST)
    copy_python()
    build_native(force)
    return True


if __name__ == "__main__":
    install(force="--force" in sys.argv)
    print("installed:", DEST)
