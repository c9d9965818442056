"""The C-ABI shared library loads and exports every symbol include/*.h declares (no compute calls: CPU-safe)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names.update(re.findall(r"\b(agr_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 8
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_every_declared_symbol(built_lib):
    from animatablegaussians_b200 import _lib, avatar, lbs, loss, lpips, optim, rasterizer, smpl_lbs, styleunet_ops  # noqa: F401  (registering modules)
    _lib.load()
    assert set(declared_symbols()) <= set(_lib.SYMBOLS), set(declared_symbols()) - set(_lib.SYMBOLS)


def test_workspace_query_and_argument_validation(built_lib):
    from animatablegaussians_b200 import _lib
    lib = _lib.load()
    ws = _lib.AgrRasterWorkspace()
    assert lib.agr_raster_workspace(1000, 2, 64, 64, 0, 5000, ctypes.byref(ws)) == _lib.AGR_OK
    assert ws.geom_bytes > 1000 * 2 * 32 and ws.binning_bytes > 5000 * 48 and ws.backward_bytes >= 1000 * 2 * 64
    assert lib.agr_raster_workspace(1000, 0, 64, 64, 0, 5000, ctypes.byref(ws)) == _lib.AGR_ERR_INVALID_ARGUMENT
    assert lib.agr_raster_workspace(1000, 33, 64, 64, 0, 5000, ctypes.byref(ws)) == _lib.AGR_ERR_INVALID_ARGUMENT
    assert lib.agr_raster_forward(None, None) == _lib.AGR_ERR_INVALID_ARGUMENT
    assert lib.agr_version().startswith(b"agr-b200")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from animatablegaussians_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_ctypes_structs_match_the_c_headers(tmp_path):
    """sizeof / offsetof of the by-value descriptor structs as gcc lays them out vs the ctypes mirrors (the host arrays
    of agr_modweight_group_* / agr_equal_linear_group_* and the rasterizer argument blocks)."""
    import subprocess
    from animatablegaussians_b200 import _lib, styleunet_ops as ops
    structs = {"AgrModWeightItem": (ops.AgrModWeightItem, "agr_styleunet.h"), "AgrEqualLinearItem": (ops.AgrEqualLinearItem, "agr_styleunet.h"),
               "AgrConvGeom": (ops.AgrConvGeom, "agr_conv.h"), "AgrConvEpilogue": (ops.AgrConvEpilogue, "agr_conv.h"),
               "AgrRasterWorkspace": (_lib.AgrRasterWorkspace, "agr_rasterizer.h"), "AgrRasterForwardArgs": (_lib.AgrRasterForwardArgs, "agr_rasterizer.h"),
               "AgrRasterBackwardArgs": (_lib.AgrRasterBackwardArgs, "agr_rasterizer.h")}
    src = ['#include <stdio.h>', '#include <stddef.h>'] + sorted({'#include "%s"' % h for _, h in structs.values()}) + ["int main(void) {"]
    for name, (cls, _) in structs.items():
        src.append('printf("%s %%zu", sizeof(%s));' % (name, name))
        for f, _t in cls._fields_:
            src.append('printf(" %%zu", offsetof(%s, %s));' % (name, f))
        src.append('printf("\\n");')
    src.append("return 0; }")
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line in filter(None, out):
        tok = line.split()
        cls = structs[tok[0]][0]
        assert int(tok[1]) == ctypes.sizeof(cls), tok[0]
        assert [int(t) for t in tok[2:]] == [getattr(cls, f).offset for f, _ in cls._fields_], tok[0]
