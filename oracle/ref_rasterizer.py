"""TEST INFRASTRUCTURE — loader for the UNMODIFIED reference CUDA rasterizer built by
oracle/build_ref.py into oracle/_ref/libref_rasterizer.so (reference sources compiled where they
lie; C shim oracle/ref_shim.cu).  Needs a GPU.  Mirrors the argument order of the reference's
pybind entry points (RAST/rasterize_points.h:18-68)."""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libref_rasterizer.so")
_lib = None


def available():
    return os.path.exists(LIB)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        lib.refrast_create.restype = C.c_void_p
        lib.refrast_destroy.argtypes = [C.c_void_p]
        lib.refrast_forward.restype = C.c_int
        lib.refrast_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float, C.c_int] + [C.c_void_p] * 4
        lib.refrast_backward.restype = None
        lib.refrast_backward.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + \
            [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 5 + [C.c_float, C.c_float] + [C.c_void_p] * 14
        lib.refrast_mark_visible.argtypes = [C.c_int] + [C.c_void_p] * 4
        lib.refrast_last_error.restype = C.c_int
        _lib = lib
    return _lib


def _p(t):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


class RefRasterizer:
    """Runs on the legacy default stream like the reference; callers must be on the default stream."""

    def __init__(self):
        self.lib = _load()
        self.h = C.c_void_p(self.lib.refrast_create())

    def __del__(self):
        try:
            self.lib.refrast_destroy(self.h)
        except Exception:
            pass

    def forward(self, bg, means3D, colors, opacities, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix,
                projmatrix, tanfovx, tanfovy, H, W, sh=None, degree=0, campos=None, prefiltered=False):
        dev = means3D.device
        self.a = dict(bg=_c(bg), means3D=_c(means3D), colors=_c(colors), opacities=_c(opacities), scales=_c(scales),
                      rotations=_c(rotations), cov3D=_c(cov3D_precomp), view=_c(viewmatrix), proj=_c(projmatrix),
                      sh=_c(sh), campos=_c(campos), mod=float(scale_modifier), tfx=float(tanfovx), tfy=float(tanfovy),
                      H=int(H), W=int(W), D=int(degree))
        a = self.a
        P = a["means3D"].shape[0]
        M = a["sh"].shape[1] if a["sh"] is not None else 0
        a["P"], a["M"] = P, M
        color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
        depth = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        alpha = torch.zeros((1, H, W), dtype=torch.float32, device=dev)
        radii = torch.zeros((P,), dtype=torch.int32, device=dev)
        R = self.lib.refrast_forward(self.h, P, a["D"], M, _p(a["bg"]), W, H, _p(a["means3D"]), _p(a["sh"]),
                                     _p(a["colors"]), _p(a["opacities"]), _p(a["scales"]), a["mod"], _p(a["rotations"]),
                                     _p(a["cov3D"]), _p(a["view"]), _p(a["proj"]), _p(a["campos"]), a["tfx"], a["tfy"],
                                     int(prefiltered), _p(color), _p(depth), _p(alpha), _p(radii))
        self.alpha, self.radii, self.num_rendered = alpha, radii, int(R)
        return color, radii, depth, alpha

    def backward(self, dL_dcolor, dL_ddepth, dL_dalpha):
        a = self.a
        P, M, H, W = a["P"], a["M"], a["H"], a["W"]
        dev = a["means3D"].device
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        g = dict(means2D=z(P, 3), conic=z(P, 2, 2), opacity=z(P, 1), colors=z(P, 3), depth=z(P, 1), means3D=z(P, 3),
                 cov3D=z(P, 6), sh=z(P, max(M, 1), 3), scales=z(P, 3), rotations=z(P, 4))
        gc, gd, ga = _c(dL_dcolor), _c(dL_ddepth), _c(dL_dalpha)
        self.lib.refrast_backward(self.h, P, a["D"], M, self.num_rendered, _p(a["bg"]), W, H, _p(a["means3D"]),
                                  _p(a["sh"]), _p(a["colors"]), _p(self.alpha), _p(a["scales"]), a["mod"],
                                  _p(a["rotations"]), _p(a["cov3D"]), _p(a["view"]), _p(a["proj"]), _p(a["campos"]),
                                  a["tfx"], a["tfy"], _p(self.radii), _p(gc), _p(gd), _p(ga), _p(g["means2D"]),
                                  _p(g["conic"]), _p(g["opacity"]), _p(g["colors"]), _p(g["depth"]), _p(g["means3D"]),
                                  _p(g["cov3D"]), _p(g["sh"]), _p(g["scales"]), _p(g["rotations"]))
        if M == 0:
            g["sh"] = z(P, 0, 3)
        return g

    def mark_visible(self, means3D, viewmatrix, projmatrix):
        m, v, p = _c(means3D), _c(viewmatrix), _c(projmatrix)
        out = torch.zeros((m.shape[0],), dtype=torch.bool, device=m.device)
        self.lib.refrast_mark_visible(m.shape[0], _p(m), _p(v), _p(p), _p(out))
        return out
