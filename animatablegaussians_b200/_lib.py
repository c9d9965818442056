"""ctypes binding of libagr_b200.so (C ABI declared in include/agr_rasterizer.h, include/agr_lbs.h, ...).

There is NO fallback: if the library is missing or a symbol is absent the import fails loudly.
Nothing under oracle/ is ever imported from here.
"""
import ctypes as C
import os

from ._build import LIB_PATH

AGR_MAX_VIEWS = 32
AGR_OK, AGR_ERR_INVALID_ARGUMENT, AGR_ERR_BINNING_CAPACITY, AGR_ERR_CUDA, AGR_ERR_WORKSPACE = range(5)

_p = C.c_void_p


class AgrRasterWorkspace(C.Structure):
    _fields_ = [("geom_bytes", C.c_size_t), ("image_bytes", C.c_size_t),
                ("binning_bytes", C.c_size_t), ("backward_bytes", C.c_size_t)]


class AgrRasterForwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int32), ("debug", C.c_int32),
        ("background", _p), ("bg_view_stride", C.c_int32),
        ("means3D", _p), ("shs", _p), ("colors_precomp", _p), ("colors_view_stride", C.c_int64),
        ("opacities", _p), ("scales", _p), ("rotations", _p), ("cov3D_precomp", _p),
        ("viewmatrix", _p), ("projmatrix", _p), ("campos", _p),
        ("tan_fovx", C.POINTER(C.c_float)), ("tan_fovy", C.POINTER(C.c_float)),
        ("out_color", _p), ("out_depth", _p), ("out_alpha", _p), ("radii", _p),
        ("geom_ws", _p), ("geom_bytes", C.c_size_t),
        ("image_ws", _p), ("image_bytes", C.c_size_t),
        ("binning_ws", _p), ("binning_bytes", C.c_size_t),
        ("capacity", C.c_int64), ("num_rendered", C.POINTER(C.c_int64)), ("device_status", _p),
    ]


class AgrRasterBackwardArgs(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("V", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32), ("scale_modifier", C.c_float), ("debug", C.c_int32),
        ("background", _p), ("bg_view_stride", C.c_int32),
        ("means3D", _p), ("shs", _p), ("colors_precomp", _p), ("colors_view_stride", C.c_int64),
        ("scales", _p), ("rotations", _p), ("cov3D_precomp", _p),
        ("viewmatrix", _p), ("projmatrix", _p), ("campos", _p),
        ("tan_fovx", C.POINTER(C.c_float)), ("tan_fovy", C.POINTER(C.c_float)),
        ("radii", _p), ("out_alpha", _p),
        ("dL_dout_color", _p), ("dL_dout_depth", _p), ("dL_dout_alpha", _p),
        ("dL_dmeans3D", _p), ("dL_dmeans2D", _p), ("dL_dcolors", _p), ("dL_dopacity", _p),
        ("dL_dcov3D", _p), ("dL_dsh", _p), ("dL_dscales", _p), ("dL_drotations", _p),
        ("geom_ws", _p), ("geom_bytes", C.c_size_t),
        ("image_ws", _p), ("image_bytes", C.c_size_t),
        ("binning_ws", _p), ("binning_bytes", C.c_size_t),
        ("backward_ws", _p), ("backward_bytes", C.c_size_t),
        ("capacity", C.c_int64), ("num_rendered", C.c_int64),
    ]


# name -> (restype, argtypes); every symbol the headers declare must be listed here
SYMBOLS = {
    "agr_raster_workspace": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                       C.POINTER(AgrRasterWorkspace)]),
    "agr_raster_forward": (C.c_int, [C.POINTER(AgrRasterForwardArgs), _p]),
    "agr_raster_backward": (C.c_int, [C.POINTER(AgrRasterBackwardArgs), _p]),
    "agr_raster_mark_visible": (C.c_int, [C.c_int32, _p, _p, _p, _p, _p]),
    "agr_last_cuda_error": (C.c_int, []),
    "agr_last_cuda_error_string": (C.c_char_p, []),
    "agr_version": (C.c_char_p, []),
}

_lib = None


def register_symbols(table):
    """Other binding modules (lbs, styleunet ops) add their C-ABI symbols here before load()."""
    SYMBOLS.update(table)
    if _lib is not None:
        _bind(_lib, table)


def _bind(lib, table):
    for name, (res, args) in table.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise ImportError("libagr_b200.so does not export %s — rebuild with "
                              "`python -m animatablegaussians_b200._build --force`" % name) from e
        fn.restype = res
        fn.argtypes = args


def load():
    """Load libagr_b200.so once. Raises ImportError (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "animatablegaussians_b200: CUDA library %s not found. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc; no GPU required). "
                "There is no CPU or PyTorch fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        _bind(lib, SYMBOLS)
        _lib = lib
    return _lib


def cuda_error_string():
    lib = load()
    return "%s (cudaError %d)" % (lib.agr_last_cuda_error_string().decode(), lib.agr_last_cuda_error())
