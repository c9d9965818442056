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
    from animatablegaussians_b200 import _lib, avatar, lbs, loss, optim, rasterizer, smpl_lbs, styleunet_ops  # noqa: F401  (registering modules)
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
