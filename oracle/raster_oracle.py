"""TEST INFRASTRUCTURE — numpy/ctypes front-end of oracle/raster_oracle.c (CPU restatement of the
reference rasterizer, RAST/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "raster_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libraster_oracle.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # FMA contraction like nvcc's default (-fmad=true) keeps the CPU restatement within ~2e-7 of the GPU
    # forward; without hardware FMA fall back to separate mul/add (~1e-6).
    flags = ["-O2", "-ffp-contract=off"]
    try:
        if " fma " in open("/proc/cpuinfo").read():
            flags = ["-O2", "-ffp-contract=fast", "-mfma"]
    except OSError:
        pass
    subprocess.check_call(["gcc", *flags, "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"])
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        lib.oracle_raster_create.restype = C.c_void_p
        lib.oracle_raster_destroy.argtypes = [C.c_void_p]
        lib.oracle_raster_forward.restype = C.c_int64
        lib.oracle_raster_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float] + [C.c_void_p] * 4
        lib.oracle_raster_backward.restype = None
        lib.oracle_raster_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float] + [C.c_void_p] * 13
        lib.oracle_raster_mark_visible.argtypes = [C.c_int] + [C.c_void_p] * 4
        _lib = lib
    return _lib


def _f(a):
    if a is None:
        return None
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RasterOracle:
    """One forward (+ optional backward) of the reference rasterizer on the CPU, fp32."""

    def __init__(self):
        self.lib = _load()
        self.h = C.c_void_p(self.lib.oracle_raster_create())

    def __del__(self):
        try:
            self.lib.oracle_raster_destroy(self.h)
        except Exception:
            pass

    def forward(self, bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                projmatrix, tanfovx, tanfovy, H, W, sh=None, degree=0, campos=None):
        self.a = dict(bg=_f(bg), means3D=_f(means3D), colors=_f(colors), opacities=_f(opacities), scales=_f(scales),
                      rotations=_f(rotations), cov3D=_f(cov3D_precomp), view=_f(viewmatrix), proj=_f(projmatrix),
                      sh=_f(sh), campos=_f(campos), mod=float(scale_modifier), tfx=float(tanfovx), tfy=float(tanfovy),
                      H=int(H), W=int(W), D=int(degree))
        a = self.a
        P = a["means3D"].shape[0]
        M = a["sh"].shape[1] if a["sh"] is not None else 0
        a["P"], a["M"] = P, M
        color = np.zeros((3, H, W), np.float32)
        depth = np.zeros((1, H, W), np.float32)
        alpha = np.zeros((1, H, W), np.float32)
        radii = np.zeros((P,), np.int32)
        R = self.lib.oracle_raster_forward(self.h, P, a["D"], M, _p(a["bg"]), W, H, _p(a["means3D"]), _p(a["sh"]),
                                           _p(a["colors"]), _p(a["opacities"]), _p(a["scales"]), a["mod"],
                                           _p(a["rotations"]), _p(a["cov3D"]), _p(a["view"]), _p(a["proj"]),
                                           _p(a["campos"]), a["tfx"], a["tfy"], _p(color), _p(depth), _p(alpha), _p(radii))
        self.alpha = alpha
        self.num_rendered = int(R)
        return color, radii, depth, alpha

    def backward(self, dL_dcolor, dL_ddepth, dL_dalpha, alpha_override=None):
        """alpha_override: forward alpha to take T_final = 1 - alpha from (backward.cu:463).  T_final of a
        saturated pixel is 1 - 0.9999..: one ulp of alpha is ~6e-4 of T_final, so gradients can only be compared
        at 1e-4 between two implementations when both start from the SAME forward alpha."""
        a = self.a
        alpha_in = self.alpha if alpha_override is None else _f(alpha_override)
        P, M, H, W = a["P"], a["M"], a["H"], a["W"]
        z = lambda *s: np.zeros(s, np.float32)
        g = dict(means2D=z(P, 3), conic=z(P, 4), opacity=z(P, 1), colors=z(P, 3), depth=z(P, 1), means3D=z(P, 3),
                 cov3D=z(P, 6), sh=z(P, max(M, 1), 3), scales=z(P, 3), rotations=z(P, 4))
        gc, gd, ga = _f(dL_dcolor), _f(dL_ddepth), _f(dL_dalpha)
        self.lib.oracle_raster_backward(self.h, P, a["D"], M, _p(a["bg"]), W, H, _p(a["means3D"]), _p(a["sh"]),
                                        _p(a["colors"]), _p(alpha_in), _p(a["scales"]), a["mod"], _p(a["rotations"]),
                                        _p(a["cov3D"]), _p(a["view"]), _p(a["proj"]), _p(a["campos"]), a["tfx"], a["tfy"],
                                        _p(gc), _p(gd), _p(ga), _p(g["means2D"]), _p(g["conic"]), _p(g["opacity"]),
                                        _p(g["colors"]), _p(g["depth"]), _p(g["means3D"]), _p(g["cov3D"]), _p(g["sh"]),
                                        _p(g["scales"]), _p(g["rotations"]))
        if M == 0:
            g["sh"] = z(P, 0, 3)
        return g


def mark_visible(means3D, viewmatrix, projmatrix):
    lib = _load()
    m, v, p = _f(means3D), _f(viewmatrix), _f(projmatrix)
    out = np.zeros((m.shape[0],), np.uint8)
    lib.oracle_raster_mark_visible(m.shape[0], _p(m), _p(v), _p(p), _p(out))
    return out.astype(bool)
